"""Numpy / cv2 restatement of the reference's constraint generation and problem assembly
(TEST INFRASTRUCTURE; checks robust_cvd_b200/host/*.cpp).  Follows SURVEY.md Appendix B:
lib/FlowConstraints.cpp:352-465 (candidates, corner score, greedy disc sampler) and
lib/PoseOptimizer.cpp:104-117, :1167-1193 (observation records).  OpenCV operators are the real
cv2 ones, so this also pins the C++ restatements of cornerMinEigenVal / cvtColor / distanceTransform.
"""
import numpy as np

f32 = np.float32


def pair_constraints(color_bgr, flow, mask, sep, inv_aspect, score=None):
    """score: optional precomputed corner response (default: the real cv2.cornerMinEigenVal, whose SIMD summation order
    differs from the C++/CUDA restatements in the last bits -- enough to swap near-equal priorities at small separations)."""
    import cv2
    h, w = mask.shape
    if score is None:
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


# ---------------------------------------------------------------------------------------------------------
# Flow-guided temporal depth filter: literal float32 restatement of DepthVideoProcessor::flowGuidedFilter
# (reference lib/Processor.cpp:315-590) and DepthVideo::project (lib/DepthVideo.cpp:637-681).  Pure-Python loops:
# small cases only.  Arrays use the local frame indexing of rcvd_flow_guided_filter (include/rcvd.h).
# ---------------------------------------------------------------------------------------------------------
def _quat_rotate(q, v):
    """Eigen::Quaternionf * Vector3f (uv = 2 q.vec x v; v + w uv + q.vec x uv), float32."""
    qv = q[:3].astype(f32); w = f32(q[3]); v = np.asarray(v, f32)
    uv = np.cross(qv, v).astype(f32); uv = (uv + uv).astype(f32)
    return (v + w * uv + np.cross(qv, uv).astype(f32)).astype(f32)


def _trunc(v):
    return int(v)   # C (int) conversion: toward zero


def flow_guided_filter(depth, cams, fwd_flow, fwd_mask, bwd_flow, bwd_mask, first_out, num_out, frame_radius, spatial_radius, median,
                       inv_aspect, far_pairs=(), far_flow=None, far_mask=None):
    """depth [F,hd,wd] f32, cams [F,9] f32 (pos xyz, quat xyzw, hFov, vFov), flows [F,h,w,2], masks [F,h,w] -> [num_out,h,w] f32."""
    F, hd, wd = depth.shape
    h, w = fwd_mask.shape[1:3] if fwd_mask is not None else far_mask.shape[1:3]
    inv_aspect = f32(inv_aspect)
    last = first_out + num_out - 1
    half = f32(0.5)

    def project(fi, ndc):   # lib/DepthVideo.cpp:655-681 with useWarp = false, then :637-653
        x = min(wd - 1, _trunc(ndc[0] * f32(wd) + half)); y = min(hd - 1, _trunc(ndc[1] / inv_aspect * f32(hd) + half))
        x = max(x, 0); y = max(y, 0)
        d = depth[fi, y, x]
        c = cams[fi].astype(f32)
        th = f32(np.tan(f32(c[7] / f32(2.0)))); tv = f32(np.tan(f32(c[8] / f32(2.0))))
        right = _quat_rotate(c[3:7], (1, 0, 0)); up = _quat_rotate(c[3:7], (0, 1, 0)); front = _quat_rotate(c[3:7], (0, 0, -1))
        rx = f32(-1.0) + f32(2.0) * ndc[0]; ry = f32(1.0) - f32(2.0) * ndc[1] / inv_aspect
        ray = (front + right * f32(rx * th) + up * f32(ry * tv)).astype(f32)
        return (c[:3] + ray * d).astype(f32)

    out = np.zeros((num_out, h, w), f32)
    far_by_src = {}
    for k, (s, d) in enumerate(far_pairs):
        far_by_src.setdefault(int(s), []).append((k, int(d)))
    for o in range(num_out):
        frame = first_out + o
        ref_pos = cams[frame, :3].astype(f32); ref_fwd = _quat_rotate(cams[frame, 3:7], (0, 0, -1))
        f0 = max(0, frame - frame_radius); f1 = min(last, frame + frame_radius)

        def sample(loc, fi):
            ndc = (f32(loc[0] / f32(w)), f32(f32(loc[1] / f32(h)) * inv_aspect))
            p = (project(fi, ndc) - ref_pos).astype(f32)
            return f32(f32(f32(p[0] * ref_fwd[0]) + f32(p[1] * ref_fwd[1])) + f32(p[2] * ref_fwd[2]))

        def step(flow, mask, loc):
            ix = min(_trunc(loc[0] + half), w - 1); iy = min(_trunc(loc[1] + half), h - 1)
            if not mask[iy, ix]:
                return None
            loc = (f32(loc[0] + flow[iy, ix, 0]), f32(loc[1] + flow[iy, ix, 1]))
            jx = _trunc(loc[0] + half); jy = _trunc(loc[1] + half)
            if jx < 0 or jx >= w or jy < 0 or jy >= h:
                return None
            return loc

        for y in range(h):
            y0, y1 = max(0, y - spatial_radius), min(h - 1, y + spatial_radius)
            for x in range(w):
                x0, x1 = max(0, x - spatial_radius), min(w - 1, x + spatial_radius)
                samples = []; ref = None
                for wy in range(y0, y1 + 1):
                    for wx in range(x0, x1 + 1):
                        samples.append(sample((f32(wx), f32(wy)), frame))
                        if wx == x and wy == y:
                            ref = samples[-1]
                        loc = (f32(wx), f32(wy))
                        for fi in range(frame + 1, f1 + 1):
                            loc = step(fwd_flow[fi - 1], fwd_mask[fi - 1], loc)
                            if loc is None:
                                break
                            samples.append(sample(loc, fi))
                        loc = (f32(wx), f32(wy))
                        for fi in range(frame - 1, f0 - 1, -1):
                            loc = step(bwd_flow[fi + 1], bwd_mask[fi + 1], loc)
                            if loc is None:
                                break
                            samples.append(sample(loc, fi))
                        for k, dst in far_by_src.get(frame, ()):
                            loc = step(far_flow[k], far_mask[k], (f32(wx), f32(wy)))
                            if loc is None:
                                break
                            samples.append(sample(loc, dst))
                with np.errstate(all="ignore"):
                    wts = [f32(np.exp(f32(-(max(s, ref) / min(s, ref)) * f32(3.0)))) for s in samples]
                    dsum = f32(0); wsum = f32(0)
                    for s, wt in zip(samples, wts):
                        dsum = f32(dsum + f32(s * wt)); wsum = f32(wsum + wt)
                    if median:
                        halfw = f32(wsum / f32(2.0)); cum = f32(0)
                        for s, wt in sorted(zip(samples, wts), key=lambda t: t[0]):
                            cum = f32(cum + wt)
                            if cum >= halfw:
                                out[o, y, x] = s
                                break
                    else:
                        out[o, y, x] = f32(dsum / wsum) if wsum > 0 else f32(0)
    return out
