"""A/B of one debug switch inside ONE process / GPU (interleaved repetitions): python tools/ab_iteration.py <hook> [reps]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from robust_cvd_b200 import solver
hook = sys.argv[1]; reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
values = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1]
spec, sc, cfg, pairs, offs, rec, med = bench.build_case("config2_300f_384x224_grid16x12_sep10")
P = solver.Problem(cfg)
P.set_frames(np.ones(cfg.num_frames, np.uint8), med); P.set_constraints(pairs, offs, rec); P.set_state(bench.initial_state(sc, cfg, P.stride))
f = getattr(P.L, hook)
res = {v: [] for v in values}
P.time_iteration(iters=2)
for r in range(reps):
    for v in values:
        assert f(P.h, C.c_int32(v)) == 0
        res[v].append(P.time_iteration(iters=5)["iter_ms"])
print(hook, {k: [round(x, 3) for x in v] for k, v in res.items()}, "means:", {k: round(float(np.mean(v)), 3) for k, v in res.items()})
