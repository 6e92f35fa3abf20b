"""Timing experiment for k_update_tma (results of the solve are INVALID in the experiment modes): serialised update-kernel time with
the DMMA warps running on resident shared memory without waiting for loads (mode 1) against the normal kernel -- separates the
cost of the inner LDS + DMMA loop from the cost of feeding it."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from robust_cvd_b200 import solver
spec, sc, cfg, pairs, offs, rec, med = bench.build_case("config2_300f_384x224_grid16x12_sep10")
P = solver.Problem(cfg)
P.set_frames(np.ones(cfg.num_frames, np.uint8), med); P.set_constraints(pairs, offs, rec); P.set_state(bench.initial_state(sc, cfg, P.stride))
P.time_iteration(iters=2)
for name, mode in (("normal", 1), ("no-load (compute only)", 1 | (1 << 8)), ("no epilogue", 1 | (2 << 8)), ("no-load, no epilogue", 1 | (3 << 8)), ("legacy cp.async", 0)):
    assert P.L.rcvd_debug_set_update_kernel(P.h, C.c_int32(mode), C.c_int32(0)) == 0
    pl = P.profile_linear(reps=3)
    big = P.profile_linear(reps=3)   # (same numbers; second call for stability)
    print(f"{name:26s} update kernels {pl['gemm_ms']:.3f} ms  = {pl['gemm_flops'] / (pl['gemm_ms'] * 1e-3) / 1e12:.2f} TFLOP/s   (potrf {pl['potrf_ms']:.3f} trsm {pl['trsm_ms']:.3f} trinv {pl['trinv_ms']:.3f} solve {pl['solve_ms']:.3f})")
