"""Numpy / cv2 restatement of the reference's constraint generation and problem assembly
(TEST INFRASTRUCTURE; checks robust_cvd_b200/host/*.cpp).  Follows SURVEY.md Appendix B:
lib/FlowConstraints.cpp:352-465 (candidates, corner score, greedy disc sampler) and
lib/PoseOptimizer.cpp:104-117, :1167-1193 (observation records).  OpenCV operators are the real
cv2 ones, so this also pins the C++ restatements of cornerMinEigenVal / cvtColor / distanceTransform.
"""
import numpy as np

f32 = np.float32


def pair_constraints(color_bgr, flow, mask, sep, inv_aspect):
    import cv2
    h, w = mask.shape
    gray = cv2.cvtColor(color_bgr, cv2.COLOR_BGR2GRAY)
    score = cv2.cornerMinEigenVal(gray, 3)
    iy, ix = np.mgrid[0:h, 0:w]
    fx1 = (ix.astype(f32) + flow[..., 0]).astype(f32); fy1 = (iy.astype(f32) + flow[..., 1]).astype(f32)
    ix1 = (fx1 + f32(0.5)).astype(np.int32); iy1 = (fy1 + f32(0.5)).astype(np.int32)      # C (int): truncation toward zero
    cand = (mask != 0) & (ix1 >= 0) & (ix1 < w) & (iy1 >= 0) & (iy1 < h)
    ys, xs = np.nonzero(cand)                                                         # row-major scan order
    order = np.argsort(-score[ys, xs], kind="stable")
    invalid = np.zeros((h, w), bool)
    dy, dx = np.mgrid[-sep:sep + 1, -sep:sep + 1]
    disk = (dx * dx + dy * dy) <= sep * sep
    sx = f32(1.0) / f32(w); sy = f32(inv_aspect) / f32(h)
    out = []
    for k in order:
        y, x = ys[k], xs[k]
        if invalid[y, x]:
            continue
        out.append((f32(x) * sx, f32(y) * sy, fx1[y, x] * sx, fy1[y, x] * sy))
        y0, y1, x0, x1 = max(0, y - sep), min(h - 1, y + sep), max(0, x - sep), min(w - 1, x + sep)
        invalid[y0:y1 + 1, x0:x1 + 1] |= disk[y0 - (y - sep):y1 - (y - sep) + 1, x0 - (x - sep):x1 - (x - sep) + 1]
    return np.array(out, f32).reshape(-1, 4), score


def observation_records(locs, depth0, depth1, inv_aspect):
    """locs [n,4] float32 (loc0.xy, loc1.xy) -> records [m,6] float32 (constraints with valid depths only)."""
    inv_aspect = f32(inv_aspect)
    rec = np.zeros((locs.shape[0], 6), f32)
    ok = np.ones(locs.shape[0], bool)
    for o, d in ((0, depth0), (1, depth1)):
        lx, ly = locs[:, 2 * o], locs[:, 2 * o + 1]
        rec[:, 3 * o] = f32(-1.0) + f32(2.0) * lx
        rec[:, 3 * o + 1] = f32(1.0) - f32(2.0) * ly / inv_aspect
        px = (lx * f32(d.shape[1])).astype(np.int32); py = (ly / inv_aspect * f32(d.shape[0])).astype(np.int32)
        px = np.clip(px, 0, d.shape[1] - 1); py = np.clip(py, 0, d.shape[0] - 1)
        sd = d[py, px]
        rec[:, 3 * o + 2] = sd
        ok &= np.isfinite(sd) & (sd > 0)
    return rec[ok]
