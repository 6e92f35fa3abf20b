"""Times the GPU constraint builder (rcvd_build_constraints, host buffers in/out) at config-2 image size against the
sequential sampler restatement (numpy/cv2, one pair) -- a side measurement for DESIGN.md, not the bench metric."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from robust_cvd_b200 import solver, synthetic
from oracle import host_ref

P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W, H, N = 384, 224, 32
sc = synthetic.Scene(N, W, H, seed=2)
rng = np.random.default_rng(0)
iy, ix = np.mgrid[0:H, 0:W]
tex = (np.sin(ix * 0.37)[..., None] * np.cos(iy * 0.23)[..., None] * 0.3 + 0.5 + rng.normal(0, 0.05, (N, H, W, 3))).astype(np.float32)
pairs = [(int(a), int((a + 1 + k % 3) % N)) for k, a in enumerate(rng.integers(0, N, P))]
flow = np.zeros((P, H, W, 2), np.float32); mask = np.zeros((P, H, W), np.uint8)
for k, (a, b) in enumerate(pairs):
    fx, fy, ok = sc.flow(a, b, ix.ravel(), iy.ravel(), rng)
    flow[k] = np.stack([fx - ix.ravel(), fy - iy.ravel()], -1).reshape(H, W, 2); mask[k] = (ok.reshape(H, W) * 255).astype(np.uint8)
for rep in range(3):
    t = time.perf_counter()
    poff, pc, _, _ = solver.build_constraints(tex, pairs, flow, mask, 10, float(sc.inv_aspect32))
    dt = time.perf_counter() - t
    print(f"GPU builder rep {rep}: {P} pairs {W}x{H}, {len(pc)} constraints, {solver.lib().rcvd_builder_last_rounds()} rounds, {dt * 1e3:.1f} ms for two passes (sizing + fill) -> {dt / 2 / P * 1e3:.3f} ms/pair")
t = time.perf_counter(); want, _ = host_ref.pair_constraints(tex[pairs[0][0]], flow[0], mask[0], 10, sc.inv_aspect32); dt = time.perf_counter() - t
print(f"numpy/cv2 sequential sampler: {dt * 1e3:.1f} ms for one pair ({len(want)} constraints)")
