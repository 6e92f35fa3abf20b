import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from robust_cvd_b200 import solver
wl = "config2_300f_384x224_grid16x12_sep10"
spec, sc, cfg, pairs, offs, rec, med = bench.build_case(wl)
P = solver.Problem(cfg)
P.set_frames(np.ones(cfg.num_frames, np.uint8), med); P.set_constraints(pairs, offs, rec); P.set_state(bench.initial_state(sc, cfg, P.stride))
for mode, name in ((1, "blocked pivot tile (default)"),):
    P.L.rcvd_debug_set_potrf_chain_warp(P.h, mode)
    t = P.time_iteration(iters=5)
    pl = P.profile_linear(reps=3)
    print(name, {k: round(v, 4) for k, v in t.items()}, {k: round(v, 4) for k, v in pl.items()} if isinstance(pl, dict) else pl)
    print("  residual", P.linear_residual(1e4))
for on in (0, 1, 150, 600, 1000000, 0, 1):
    P.set_fused_substitution(on)
    t = P.time_iteration(iters=5); pl = P.profile_linear(reps=3)
    print("fused substitution", on, {k: round(v, 4) for k, v in t.items()}, "solve_ms", round(pl["solve_ms"], 4), "residual", P.linear_residual(1e4)["rel_residual"])
