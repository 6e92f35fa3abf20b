"""In-process A/B of the factorisation kernels' variants (config 2).  Modes marked (*) leave INVALID numbers in the factor (timing only)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from robust_cvd_b200 import solver
spec, sc, cfg, pairs, offs, rec, med = bench.build_case("config2_300f_384x224_grid16x12_sep10")
P = solver.Problem(cfg)
P.set_frames(np.ones(cfg.num_frames, np.uint8), med); P.set_constraints(pairs, offs, rec); P.set_state(bench.initial_state(sc, cfg, P.stride))
P.time_iteration(iters=2)
def show(name):
    pl = P.profile_linear(reps=3)
    it = P.time_iteration(iters=5)
    print(f"{name:46s} iter {it['iter_ms']:.3f} ms | serial: update {pl['gemm_ms']:.3f} ms = {pl['gemm_flops'] / (pl['gemm_ms'] * 1e-3) / 1e12:.2f} TF/s, potrf {pl['potrf_ms']:.3f}, trsm {pl['trsm_ms']:.3f}, trinv {pl['trinv_ms']:.3f}, solve {pl['solve_ms']:.3f}", flush=True)
def upd(tma=1, dbg=0, team=None, ipc=0, reserve=0):
    v = tma | (dbg << 8) | (((team + 1) << 16) if team is not None else 0)
    assert P.L.rcvd_debug_set_update_kernel(P.h, C.c_int32(v), C.c_int32(ipc | (reserve << 16))) == 0
upd(); show("default (two-team shape up to 1 item/SM)")
for ipc in (1, 2, 4):
    upd(ipc=ipc); show(f"4-warp launches: at most {ipc} items per CTA")
for rsv in (4, 8, 16):
    upd(reserve=rsv); show(f"narrow-level overlapped updates leave {rsv} SMs free")
upd(reserve=4, ipc=2); show("reserve 4 + ipc 2")
upd(team=0); show("update: 4-warp shape only")
upd(team=2); show("update: two-team shape up to 2 items/SM")
upd(team=1000); show("update: two-team shape always")
upd()
P.set_trsm_ll(3); show("trsm: no deep prefetch (round-1 staging)"); P.set_trsm_ll(1)
upd(dbg=1); show("(*) update without loads")
upd(dbg=2); show("(*) update without epilogue")
upd(dbg=3); show("(*) update without loads and epilogue")
upd(tma=0); show("legacy cp.async update kernel")
