"""2-GPU test of the pair-sharded path: NCCL all-reduce of H / g / cost inside the library must reproduce
the single-GPU evaluation and LM result (skipped on boxes with one GPU)."""
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
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sc, cfg, pairs, offs, rec, med = helpers.make_case(num_frames=8, depth_type=abi.DEPTH_GRID, depth_grid_x=4, depth_grid_y=4)
x = helpers.initial_state(sc, cfg, 23, 7, 16)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0: uid = torch.from_numpy(solver.nccl_unique_id()).cuda()
dist.broadcast(uid, 0)
P = solver.Problem(cfg, device=local)
P.init_comm(world, rank, uid.cpu().numpy()); P.set_structure(pairs)
sel = sharding.lpt_partition(np.diff(offs), world)[rank]
lp, lo, lr = sharding.take_pairs(pairs, offs, rec, sel)
P.set_frames(np.ones(8, np.uint8), med); P.set_constraints(lp, lo, lr); P.set_state(x)
c, g = P.evaluate(True)
s = P.solve(abi.default_solve_options(max_iterations=40))
xs = P.get_state()
if rank == 0:
    json.dump({"cost": c, "grad": g.tolist(), "final": s.final_cost, "iters": s.iterations, "x": xs.ravel().tolist()}, open(os.environ["RCVD_OUT"], "w"))
dist.destroy_process_group()
'''


def test_two_gpu_sharded_matches_single(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from robust_cvd_b200 import abi, solver
    from tests import helpers
    script = tmp_path / "worker.py"; script.write_text(WORKER)
    out = tmp_path / "out.json"
    env = dict(os.environ, RCVD_ROOT=ROOT, RCVD_OUT=str(out))
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                           "--master-port", "29611", str(script)], env=env, timeout=300)
    r = json.load(open(out))
    sc, cfg, pairs, offs, rec, med = helpers.make_case(num_frames=8, depth_type=abi.DEPTH_GRID, depth_grid_x=4, depth_grid_y=4)
    G = solver.Problem(cfg)
    helpers.setup_problem(G, cfg, pairs, offs, rec, med, helpers.initial_state(sc, cfg, 23, 7, 16))
    c, g = G.evaluate(True)
    s = G.solve(abi.default_solve_options(max_iterations=40))
    assert abs(r["cost"] - c) <= 1e-11 * abs(c)
    assert np.abs(np.array(r["grad"]) - g).max() <= 1e-9 * max(1.0, np.abs(g).max())
    assert abs(r["final"] - s.final_cost) <= 1e-6 * s.final_cost and abs(r["iters"] - s.iterations) <= 2
    assert np.linalg.norm(np.array(r["x"]) - G.get_state().ravel()) <= 1e-4 * np.linalg.norm(G.get_state())
