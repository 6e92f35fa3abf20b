"""Where the per-call overhead of the C-ABI path goes (create / upload+structure / solve / download / destroy)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from robust_cvd_b200 import abi, solver
spec, sc, cfg, pairs, offs, rec, med = bench.build_case("config2_300f_384x224_grid16x12_sep10")
x0 = bench.initial_state(sc, cfg, solver.frame_stride(cfg))
opt = abi.default_solve_options(max_iterations=int(sys.argv[1]) if len(sys.argv) > 1 else 20)
opt.function_tolerance = 0.0; opt.parameter_tolerance = 0.0; opt.gradient_tolerance = 0.0
for rep in range(3):
    t = [time.perf_counter()]
    Q = solver.Problem(cfg); t.append(time.perf_counter())
    Q.set_frames(np.ones(cfg.num_frames, np.uint8), med); Q.set_constraints(pairs, offs, rec); Q.set_state(x0); t.append(time.perf_counter())
    c = Q.evaluate(); t.append(time.perf_counter())          # forces structure build + uploads + one cost evaluation
    s = Q.solve(opt); t.append(time.perf_counter())
    xs = Q.get_state(); t.append(time.perf_counter())
    Q.close(); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print(f"rep {rep}: create {d[0]:.1f} ms, set {d[1]:.1f}, first evaluate (structure+alloc+upload) {d[2]:.1f}, solve {d[3]:.1f} ({s.iterations} it, device eval {s.eval_ms:.1f} lin {s.linear_ms:.1f} cost {s.cost_ms:.1f}), get {d[4]:.1f}, destroy {d[5]:.1f}, total {d.sum():.1f}")
