"""Per-level kernel times of one factorisation (config 2): serialised on one stream, and the main-stream view of the two-stream schedule
(time between consecutive main-stream launches: shows what the chain of a level costs with the overlapped work running beside it)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from robust_cvd_b200 import solver
wl = sys.argv[1] if len(sys.argv) > 1 else "config2_300f_384x224_grid16x12_sep10"
spec, sc, cfg, pairs, offs, rec, med = bench.build_case(wl)
P = solver.Problem(cfg)
P.set_frames(np.ones(cfg.num_frames, np.uint8), med); P.set_constraints(pairs, offs, rec); P.set_state(bench.initial_state(sc, cfg, P.stride))
print(P.time_iteration(iters=3))
for reps, name in ((3, "serialised"), (-3, "two-stream, main-stream view")):
    pl = P.profile_linear(reps=reps)
    out = (C.c_double * (64 * 6))()
    n = P.L.rcvd_debug_level_profile(P.h, out, 64)
    a = np.array(list(out)[:n * 6]).reshape(n, 6) * 1e3
    print(f"--- {name}: per level us [potrf trinv trsm update]   (sums ms: potrf {a[:,1].sum()/1e3:.2f} trinv {a[:,2].sum()/1e3:.2f} trsm {a[:,3].sum()/1e3:.2f} update {a[:,4].sum()/1e3:.2f})")
    for l in range(n):
        print(f"{l:3d}  {a[l,1]:7.1f} {a[l,2]:7.1f} {a[l,3]:7.1f} {a[l,4]:8.1f}")
