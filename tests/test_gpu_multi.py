"""Multi-GPU tests (NCCL, one process per GPU; skipped when the box has fewer GPUs than ranks).

Pair-sharded accumulation + distributed factorisation: every rank accumulates its pair shard, H is summed
onto the owners of its blocks, the wide early levels of the block Cholesky are factored by the owners of the frames with one fused
NCCL broadcast of the new factor blocks per level, the narrow tail is replicated.  Everything observable must equal the single-GPU
solve: cost, gradient, the step of the damped system (device-side residual), the LM trajectory.  The round-1 scheme (all-reduce of H,
replicated factorisation) stays selectable and is tested too.  Reference semantics: lib/PoseOptimizer.cpp:954-987."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
sys.path.insert(0, os.environ["RCVD_ROOT"])
import numpy as np, torch, torch.distributed as dist
from robust_cvd_b200 import abi, solver, sharding
from tests import helpers
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
frames, gx, gy, distributed = int(os.environ["RCVD_FRAMES"]), int(os.environ["RCVD_GX"]), int(os.environ["RCVD_GY"]), int(os.environ["RCVD_DIST"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sc, cfg, pairs, offs, rec, med = helpers.make_case(num_frames=frames, depth_type=abi.DEPTH_GRID, depth_grid_x=gx, depth_grid_y=gy)
off_d, nd = helpers.layout_numbers(cfg)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0: uid = torch.from_numpy(solver.nccl_unique_id()).cuda()
dist.broadcast(uid, 0)
P = solver.Problem(cfg, device=local)
x = helpers.initial_state(sc, cfg, P.stride, off_d, nd)
P.init_comm(world, rank, uid.cpu().numpy()); P.set_structure(pairs); P.set_distributed(bool(distributed))
sel = sharding.lpt_partition(np.diff(offs), world)[rank]
lp, lo, lr = sharding.take_pairs(pairs, offs, rec, sel)
P.set_frames(np.ones(frames, np.uint8), med); P.set_constraints(lp, lo, lr); P.set_state(x)
info = P.distribution_info()
c, g = P.evaluate(True)
res = [P.linear_residual(r) for r in (1e4, 1e8)]
opt = abi.default_solve_options(max_iterations=40)
s = P.solve(opt)
xs = P.get_state()
tm = P.time_iteration(iters=2)          # exercises the captured graph (NCCL broadcasts inside) after the LM loop's warm-up
c2 = P.evaluate()
json.dump({"cost": c, "grad": g.tolist(), "final": s.final_cost, "iters": s.iterations, "x": xs.ravel().tolist(), "info": info, "res": res, "cost_after": c2, "iter_ms": tm["iter_ms"]},
          open(os.environ["RCVD_OUT"] + f".{rank}", "w"))
dist.destroy_process_group()
'''


def _run(tmp_path, world, frames, gx, gy, distributed, port):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from robust_cvd_b200 import abi, solver
    from tests import helpers
    script = tmp_path / "worker.py"; script.write_text(WORKER)
    out = tmp_path / "out.json"
    env = dict(os.environ, RCVD_ROOT=ROOT, RCVD_OUT=str(out), RCVD_FRAMES=str(frames), RCVD_GX=str(gx), RCVD_GY=str(gy), RCVD_DIST=str(distributed))
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                           "--master-port", str(port), str(script)], env=env, timeout=600)
    rs = [json.load(open(f"{out}.{r}")) for r in range(world)]
    sc, cfg, pairs, offs, rec, med = helpers.make_case(num_frames=frames, depth_type=abi.DEPTH_GRID, depth_grid_x=gx, depth_grid_y=gy)
    off_d, nd = helpers.layout_numbers(cfg)
    G = solver.Problem(cfg)
    helpers.setup_problem(G, cfg, pairs, offs, rec, med, helpers.initial_state(sc, cfg, G.stride, off_d, nd))
    c, g = G.evaluate(True)
    s = G.solve(abi.default_solve_options(max_iterations=40))
    xg = G.get_state()
    for r in rs:      # every rank holds the same, correct, answer
        assert r["info"]["distributed"] == distributed
        assert abs(r["cost"] - c) <= 1e-11 * abs(c)
        assert np.abs(np.array(r["grad"]) - g).max() <= 1e-9 * max(1.0, np.abs(g).max())
        for q in r["res"]:
            assert q["pivot_fail"] == 0 and q["rel_residual"] < 1e-8, q
        assert abs(r["final"] - s.final_cost) <= 1e-6 * s.final_cost and abs(r["iters"] - s.iterations) <= 2
        assert np.linalg.norm(np.array(r["x"]).reshape(xg.shape) - xg) <= 1e-4 * np.linalg.norm(xg)
        assert abs(r["cost_after"] - r["final"]) <= 1e-9 * abs(r["final"])
    if distributed:
        assert 0 < rs[0]["info"]["first_replicated_level"] <= rs[0]["info"]["levels"]
        assert sum(r["info"]["frames_owned"] for r in rs) == frames       # the ranks partition the frames
    return rs


def test_two_gpu_sharded_matches_single(tmp_path):
    _run(tmp_path, 2, 8, 4, 4, 0, 29611)                  # round-1 scheme: all-reduce of H, replicated factorisation


def test_two_gpu_distributed_factorisation_matches_single(tmp_path):
    _run(tmp_path, 2, 40, 4, 4, 1, 29612)


def test_four_gpu_distributed_factorisation_matches_single(tmp_path):
    _run(tmp_path, 4, 48, 5, 3, 1, 29613)


def test_two_gpu_distributed_large_blocks(tmp_path):
    _run(tmp_path, 2, 24, 20, 14, 1, 29614)               # npad > 224: panel/trailing potrf and inverse-times-block TRSM, distributed
