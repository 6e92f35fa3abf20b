"""N>1 host logic on CPU: world_size-2 gloo run of the pair-sharded accumulation -- each rank evaluates
its LPT shard (the oracle stands in for the device kernels here), one all-reduce restores the full
gradient / normal matrix / cost, exactly what rcvd_problem_init_comm does with NCCL on the GPUs."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from robust_cvd_b200 import abi, sharding
    from oracle import oracle
    from tests import helpers
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle.set_threads(1)
    sc, cfg, pairs, offs, rec, med = helpers.make_case(depth_type=abi.DEPTH_GRID, depth_grid_x=4, depth_grid_y=4)
    off_d, nd = helpers.layout_numbers(cfg)
    sel = sharding.lpt_partition(np.diff(offs), world)[rank]
    lp_, lo, lr = sharding.take_pairs(pairs, offs, rec, sel)
    # static part on the shard, no regularisers (they belong to rank f % world; here: add them on rank 0 only via a second problem)
    cfg_s = abi.default_config(cfg.num_frames, cfg.aspect, depth_type=abi.DEPTH_GRID, depth_grid_x=4, depth_grid_y=4,
                               scale_reg=0.0, depth_deform_reg=0.0, spatial_deform_reg=0.0, focal_reg=0.0)
    O = oracle.OracleProblem(cfg_s)
    x = helpers.initial_state(sc, cfg, O.stride, off_d, nd)
    helpers.setup_problem(O, cfg_s, lp_, lo, lr, med, x)
    c, g = O.evaluate(True); H = O.normal_matrix_dense()
    if rank == 0:    # regulariser rows
        cfg_r = abi.default_config(cfg.num_frames, cfg.aspect, depth_type=abi.DEPTH_GRID, depth_grid_x=4, depth_grid_y=4)
        R = oracle.OracleProblem(cfg_r)
        helpers.setup_problem(R, cfg_r, np.zeros((0, 2), np.int32), np.zeros(1, np.int64), np.zeros((0, 6), np.float32), med, x)
        cr, gr = R.evaluate(True); c += cr; g = g + gr; H = H + R.normal_matrix_dense()
    buf = torch.from_numpy(np.concatenate([[c], g, H.ravel()]))
    dist.all_reduce(buf)
    if rank == 0:
        F = oracle.OracleProblem(cfg)
        helpers.setup_problem(F, cfg, pairs, offs, rec, med, x)
        cf, gf = F.evaluate(True); Hf = F.normal_matrix_dense()
        out = buf.numpy()
        q.put((abs(out[0] - cf) / cf, np.abs(out[1:1 + gf.size] - gf).max(), np.abs(out[1 + gf.size:].reshape(Hf.shape) - Hf).max() / np.abs(Hf).max(),
               [int(len(s)) for s in sharding.lpt_partition(np.diff(offs), world)]))
    dist.destroy_process_group()


def test_pair_sharded_accumulation_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = q.get(timeout=240)
    for p in procs: p.join(timeout=60)
    assert res[0] < 1e-13 and res[1] < 1e-9 and res[2] < 1e-12, res
    assert sum(res[3]) == 30 and min(res[3]) >= 10


def test_lpt_partition_balances_and_covers():
    from robust_cvd_b200 import sharding
    rng = np.random.default_rng(0)
    counts = rng.integers(100, 700, 1766)
    for n in (1, 2, 4, 8):
        parts = sharding.lpt_partition(counts, n)
        allidx = np.sort(np.concatenate(parts))
        np.testing.assert_array_equal(allidx, np.arange(1766))
        loads = np.array([counts[p].sum() for p in parts])
        assert loads.max() - loads.min() <= counts.max()
