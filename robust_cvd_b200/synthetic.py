"""Seeded synthetic scenes for tests and bench.

Array-level generator: a convex "room" of planes seen by a smoothly moving
camera, exact flow by reprojection (+ sub-pixel noise), and an emulated
"network" depth = GT depth x smooth per-frame scale field x noise (MiDaS / RAFT
weights are not in this image, so their outputs are emulated).  Conventions
follow the reference: camera looks along -z, NDC x in [-1,1] left->right, y in
[-1,1] bottom->top (lib/PoseOptimizer.cpp:104-106, :175-221), per-frame
parameters [t(3), angle-axis(3), tan(vFov/2)].

The constraint records produced here are exactly what
`rcvd_problem_set_constraints` consumes (float32 ndc0.xy, depth0, ndc1.xy,
depth1; Observation ctor lib/PoseOptimizer.cpp:104-117).
"""
import numpy as np


def hierarchical2_pairs(num_frames, two_way=True):
    """Frame pairs of the reference's HIERARCHICAL2 schedule (utils/frame_sampling.py:77-120)."""
    pairs = set()
    if num_frames < 2:
        return []
    max_level = int(np.floor(np.log2(num_frames - 1)))
    for level in range(0, max_level + 1):
        dist = 1 << level
        step = 1 << max(0, level - 1)
        for start in range(0, num_frames, step):
            for sign in ((-1, 1) if two_way else (1,)):
                end = start + sign * dist
                if 0 <= end < num_frames:
                    pairs.add((start, end))
    return sorted(pairs)


def rodrigues(w):
    """Rotation matrix of an angle-axis vector (exact exponential map)."""
    w = np.asarray(w, np.float64)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


class Scene:
    def __init__(self, num_frames, width, height, seed=0, motion=0.02, rot_deg=0.3,
                 scale_sigma=0.15, depth_noise=0.01, flow_noise=0.1, focal_long=0.3461538376301239,
                 dolly=0.0, hole_fraction=0.0):
        """dolly: constant speed along the initial viewing direction (-z), scene units per frame -- a divergent (expanding)
        flow field; hole_fraction: fraction of every frame covered by seeded elliptic occluders (dynamic-mask holes)."""
        rng = np.random.default_rng(seed)
        self.N, self.w, self.h = num_frames, width, height
        self.aspect32 = np.float32(width) / np.float32(height)   # lib/DepthVideo.cpp:113
        self.inv_aspect32 = np.float32(1.0) / self.aspect32
        self.aspect = float(self.aspect32)
        self.phi = focal_long / self.aspect if self.aspect >= 1.0 else focal_long
        self.flow_noise, self.depth_noise = flow_noise, depth_noise
        # smooth random walk of the camera
        vel = np.cumsum(rng.normal(0, motion * 0.3, (num_frames, 3)), axis=0) * 0.2 + rng.normal(0, motion, (num_frames, 3))
        self.t = np.cumsum(vel, axis=0); self.t -= self.t[0]
        if dolly:
            self.t[:, 2] -= dolly * np.arange(num_frames)
        self.hole_fraction = float(hole_fraction)
        wv = np.cumsum(rng.normal(0, np.deg2rad(rot_deg), (num_frames, 3)), axis=0)
        self.w_aa = wv - wv[0]
        self.R = np.stack([rodrigues(v) for v in self.w_aa])
        # convex room: planes n.X = d with the camera inside (n.X < d for all planes)
        self.planes = [
            (np.array([0.0, 0.0, -1.0]), 4.0),     # back wall at z = -4
            (np.array([0.0, -1.0, 0.0]), 1.2),     # floor  y = -1.2
            (np.array([-1.0, 0.0, -0.15]), 3.0),   # left wall
            (np.array([1.0, 0.0, -0.25]), 3.5),    # right wall
            (np.array([0.0, 1.0, -0.2]), 2.5),     # ceiling
        ]
        # per-frame smooth multiplicative field of the emulated network depth (5x3 lattice, log-normal)
        self.field = np.exp(rng.normal(0, scale_sigma, (num_frames, 3, 5)))
        self.global_scale = np.exp(rng.normal(0.0, 0.3, num_frames))   # per-frame unknown scale
        self.seed = seed

    # --- geometry ---------------------------------------------------------
    def ray_depth(self, f, x, y):
        """GT z-depth along the rays of frame f through NDC points (x, y)."""
        d = np.stack([x * self.phi * self.aspect, y * self.phi, -np.ones_like(x)], axis=-1) @ self.R[f].T
        best = np.full(x.shape, np.inf)
        for n, dist in self.planes:
            denom = d @ n
            s = (dist - self.t[f] @ n) / np.where(np.abs(denom) < 1e-12, 1e-12, denom)
            s = np.where((denom > 1e-9) & (s > 0), s, np.inf)
            best = np.minimum(best, s)
        return best

    def net_field(self, f, x, y):
        """Smooth multiplicative error of the emulated network depth at NDC (x,y)."""
        F = self.field[f]
        gx = (x + 1) * 0.5 * (F.shape[1] - 1); gy = (y + 1) * 0.5 * (F.shape[0] - 1)
        ix = np.clip(np.floor(gx).astype(int), 0, F.shape[1] - 2); iy = np.clip(np.floor(gy).astype(int), 0, F.shape[0] - 2)
        rx = gx - ix; ry = gy - iy
        v = (F[iy, ix] * (1 - rx) * (1 - ry) + F[iy, ix + 1] * rx * (1 - ry) + F[iy + 1, ix] * (1 - rx) * ry + F[iy + 1, ix + 1] * rx * ry)
        return v * self.global_scale[f]

    def pixel_hash_noise(self, f, px, py):
        """Deterministic per-pixel noise so that a 'depth image' is a function of (frame,pixel)."""
        hsh = (px.astype(np.uint64) * np.uint64(73856093)) ^ (py.astype(np.uint64) * np.uint64(19349663)) ^ np.uint64((f + 1) * 83492791 + self.seed)
        hsh = (hsh * np.uint64(6364136223846793005) + np.uint64(1442695040888963407)) >> np.uint64(33)
        u = (hsh.astype(np.float64) / float(1 << 31)) - 0.5
        return 1.0 + self.depth_noise * 3.4641 * u   # uniform with std = depth_noise

    def net_depth_at_pixel(self, f, px, py):
        """Emulated network depth image value (float32) at integer pixel (px, py) of frame f."""
        x = -1.0 + 2.0 * px / self.w; y = 1.0 - 2.0 * py / self.h
        d = self.ray_depth(f, x, y) * self.net_field(f, x, y) * self.pixel_hash_noise(f, px, py)
        disp = (1.0 / d).astype(np.float32)
        return (np.float32(1.0) / disp).astype(np.float32)      # stored as disparity, loaded as 1/disp (lib/DepthStream.cpp:200-211)

    def depth_image(self, f):
        py, px = np.mgrid[0:self.h, 0:self.w]
        return self.net_depth_at_pixel(f, px.ravel(), py.ravel()).reshape(self.h, self.w)

    def flow(self, a, b, ix, iy, rng=None):
        """Exact flow a->b for integer source pixels (+ noise): returns target sub-pixel fx1, fy1 (float32)."""
        x = -1.0 + 2.0 * ix / self.w; y = 1.0 - 2.0 * iy / self.h
        D = self.ray_depth(a, x, y)
        d = np.stack([x * self.phi * self.aspect, y * self.phi, -np.ones_like(x)], axis=-1) @ self.R[a].T
        X = self.t[a] + d * D[..., None]
        q = (X - self.t[b]) @ self.R[b]
        depth = -q[..., 2]
        pxn = q[..., 0] / depth / (self.phi * self.aspect); pyn = q[..., 1] / depth / self.phi
        fx1 = (pxn + 1.0) * 0.5 * self.w; fy1 = (1.0 - pyn) * 0.5 * self.h
        if rng is not None and self.flow_noise > 0:
            fx1 = fx1 + rng.normal(0, self.flow_noise, fx1.shape); fy1 = fy1 + rng.normal(0, self.flow_noise, fy1.shape)
        ok = depth > 1e-3
        return fx1.astype(np.float32), fy1.astype(np.float32), ok

    # --- dynamic-mask holes (config 5) ---------------------------------------
    def holes(self, f):
        """Seeded elliptic occluders of frame f as rows (cx, cy, ax, ay) in pixels; their union covers ~hole_fraction of the
        image (emulates dynamic-object masks: constraints starting or ending inside are dropped)."""
        if self.hole_fraction <= 0:
            return np.zeros((0, 4))
        rng = np.random.default_rng(self.seed * 7919 + 31 * f + 5)
        n = 6
        area = -np.log(1.0 - min(self.hole_fraction, 0.95)) * self.w * self.h / n      # Poisson coverage: 1 - exp(-n a / A)
        ax = np.sqrt(area / np.pi) * rng.uniform(0.7, 1.4, n); ay = area / (np.pi * ax)
        return np.stack([rng.uniform(0, self.w, n), rng.uniform(0, self.h, n), ax, ay], axis=1)

    def visible(self, f, px, py):
        """False where pixel (px, py) of frame f lies inside one of its holes."""
        ok = np.ones(np.shape(px), bool)
        for cx, cy, ax, ay in self.holes(f):
            ok &= ((px - cx) / ax) ** 2 + ((py - cy) / ay) ** 2 > 1.0
        return ok

    def mask_ratio(self, a, b, stride=8):
        """Emulated flow-mask ratio of the pair (flow.py:49-66): share of (lattice) pixels of a whose flow target is a visible
        pixel of b, minimum over the two directions -- the score column of flow_list.json."""
        out = 1.0
        py, px = np.mgrid[0:self.h:stride, 0:self.w:stride]
        px = px.ravel(); py = py.ravel()
        for s, d in ((a, b), (b, a)):
            fx, fy, ok = self.flow(s, d, px, py)
            ix = (fx + np.float32(0.5)).astype(np.int64); iy = (fy + np.float32(0.5)).astype(np.int64)
            ok = ok & (fx >= 0) & (fy >= 0) & (ix < self.w) & (iy < self.h) & self.visible(s, px, py)
            ok[ok] &= self.visible(d, ix[ok], iy[ok])
            out = min(out, float(np.mean(ok)))
        return out

    def filtered_pairs(self, min_mask_ratio, pairs=None):
        """Pairs kept by the dataset's overlap filter `score > min_mask_ratio` (loaders/video_dataset.py:129-136)."""
        pairs = hierarchical2_pairs(self.N) if pairs is None else pairs
        score = {}
        keep = []
        for (a, b) in pairs:
            k = (min(a, b), max(a, b))
            if k not in score:
                score[k] = self.mask_ratio(*k)
            if score[k] > min_mask_ratio:
                keep.append((a, b))
        return keep

    # --- constraint records -----------------------------------------------
    def pair_records(self, a, b, sep, rng, valid_fraction=1.0):
        """Records of directed pair a->b, emulating the greedy disc sampler's density
        (lib/FlowConstraints.cpp:352-397): jittered lattice with >= sep px separation for sep > 0,
        every pixel for sep == 0.  Follows Appendix B of SURVEY.md for the float32 arithmetic."""
        w, h = self.w, self.h
        if sep <= 0:
            iy, ix = np.mgrid[0:h, 0:w]; ix = ix.ravel(); iy = iy.ravel()
        else:
            cell = sep * 1.2 + 0.04
            nx = max(int(w / cell), 1); ny = max(int(h / cell), 1)
            gy, gx = np.mgrid[0:ny, 0:nx]
            jit = 0.5 * (cell - sep)
            ix = np.floor((gx + 0.5) * (w / nx) + rng.uniform(-jit, jit, gx.shape)).astype(int).ravel()
            iy = np.floor((gy + 0.5) * (h / ny) + rng.uniform(-jit, jit, gy.shape)).astype(int).ravel()
            ix = np.clip(ix, 0, w - 1); iy = np.clip(iy, 0, h - 1)
        if valid_fraction < 1.0:
            keep = rng.uniform(size=ix.shape) < valid_fraction
            ix, iy = ix[keep], iy[keep]
        fx1, fy1, ok = self.flow(a, b, ix, iy, rng)
        ix1 = (fx1 + np.float32(0.5)).astype(np.int32); iy1 = (fy1 + np.float32(0.5)).astype(np.int32)   # C (int) truncation
        ok &= (ix1 >= 0) & (ix1 < w) & (iy1 >= 0) & (iy1 < h) & (fx1 >= 0) & (fy1 >= 0)
        if self.hole_fraction > 0:
            ok &= self.visible(a, ix, iy)
            ok[ok] &= self.visible(b, ix1[ok], iy1[ok])
        ix, iy, fx1, fy1 = ix[ok], iy[ok], fx1[ok], fy1[ok]
        sx = np.float32(1.0) / np.float32(w); sy = self.inv_aspect32 / np.float32(h)
        loc0x = ix.astype(np.float32) * sx; loc0y = iy.astype(np.float32) * sy
        loc1x = fx1 * sx; loc1y = fy1 * sy
        rec = np.empty((ix.size, 6), np.float32)
        two = np.float32(2.0)
        rec[:, 0] = np.float32(-1.0) + two * loc0x
        rec[:, 1] = np.float32(1.0) - two * loc0y / self.inv_aspect32
        rec[:, 3] = np.float32(-1.0) + two * loc1x
        rec[:, 4] = np.float32(1.0) - two * loc1y / self.inv_aspect32
        px0 = (loc0x * np.float32(w)).astype(np.int32); py0 = (loc0y / self.inv_aspect32 * np.float32(h)).astype(np.int32)
        px1 = (loc1x * np.float32(w)).astype(np.int32); py1 = (loc1y / self.inv_aspect32 * np.float32(h)).astype(np.int32)
        px1 = np.clip(px1, 0, w - 1); py1 = np.clip(py1, 0, h - 1)
        rec[:, 2] = self.net_depth_at_pixel(a, px0, py0)
        rec[:, 5] = self.net_depth_at_pixel(b, px1, py1)
        good = np.isfinite(rec[:, 2]) & (rec[:, 2] > 0) & np.isfinite(rec[:, 5]) & (rec[:, 5] > 0)
        return rec[good]

    def constraints(self, pairs=None, sep=10, valid_fraction=1.0):
        rng = np.random.default_rng(self.seed + 12345)
        if pairs is None:
            pairs = hierarchical2_pairs(self.N)
        recs, offs = [], [0]
        for (a, b) in pairs:
            r = self.pair_records(a, b, sep, rng, valid_fraction)
            recs.append(r); offs.append(offs[-1] + r.shape[0])
        return (np.asarray(pairs, np.int32).reshape(-1, 2), np.asarray(offs, np.int64),
                np.concatenate(recs, axis=0) if recs else np.zeros((0, 6), np.float32))

    def triplet_records(self, f, sep, rng, w_static=1.0, w_dynamic=0.5, dynamic_fraction=0.2):
        """Scene-flow smoothness constraints centred on frame f (f-1, f, f+1), emulating the triplet sampler
        (lib/FlowConstraints.cpp:467-550): records [n][10] = 3 x (ndc.xy, depth) + weight."""
        w, h = self.w, self.h
        cell = max(sep, 1) * 1.2 + 0.04
        nx = max(int(w / cell), 1); ny = max(int(h / cell), 1)
        gy, gx = np.mgrid[0:ny, 0:nx]
        ix = np.clip(np.floor((gx + 0.5) * (w / nx)).astype(int).ravel(), 0, w - 1); iy = np.clip(np.floor((gy + 0.5) * (h / ny)).astype(int).ravel(), 0, h - 1)
        fx0, fy0, ok0 = self.flow(f, f - 1, ix, iy, rng); fx2, fy2, ok2 = self.flow(f, f + 1, ix, iy, rng)
        ok = ok0 & ok2
        for fx, fy in ((fx0, fy0), (fx2, fy2)):
            ok &= (fx >= 0) & (fy >= 0) & ((fx + np.float32(0.5)).astype(np.int32) < w) & ((fy + np.float32(0.5)).astype(np.int32) < h)
        ix, iy, fx0, fy0, fx2, fy2 = ix[ok], iy[ok], fx0[ok], fy0[ok], fx2[ok], fy2[ok]
        sx = np.float32(1.0) / np.float32(w); sy = self.inv_aspect32 / np.float32(h)
        rec = np.empty((ix.size, 10), np.float32)
        for o, (frame, lx, ly) in enumerate(((f - 1, fx0 * sx, fy0 * sy), (f, ix.astype(np.float32) * sx, iy.astype(np.float32) * sy), (f + 1, fx2 * sx, fy2 * sy))):
            rec[:, 3 * o] = np.float32(-1.0) + np.float32(2.0) * lx
            rec[:, 3 * o + 1] = np.float32(1.0) - np.float32(2.0) * ly / self.inv_aspect32
            px = np.clip((lx * np.float32(w)).astype(np.int32), 0, w - 1); py = np.clip((ly / self.inv_aspect32 * np.float32(h)).astype(np.int32), 0, h - 1)
            rec[:, 3 * o + 2] = self.net_depth_at_pixel(frame, px, py)
        rec[:, 9] = np.where(rng.uniform(size=ix.size) < dynamic_fraction, w_dynamic, w_static).astype(np.float32)
        return rec

    def triplets(self, sep=10, **kw):
        rng = np.random.default_rng(self.seed + 4242)
        centers, offs, recs = [], [0], []
        for f in range(1, self.N - 1):
            r = self.triplet_records(f, sep, rng, **kw)
            centers.append(f); recs.append(r); offs.append(offs[-1] + r.shape[0])
        return np.asarray(centers, np.int32), np.asarray(offs, np.int64), (np.concatenate(recs) if recs else np.zeros((0, 10), np.float32))

    def median_depths(self, stride=4):
        """Median over the emulated depth image (subsampled lattice keeps it cheap; a statistic only)."""
        out = np.zeros(self.N)
        py, px = np.mgrid[0:self.h:stride, 0:self.w:stride]
        for f in range(self.N):
            d = self.net_depth_at_pixel(f, px.ravel(), py.ravel())
            d = np.sort(d); out[f] = float(d[d.size // 2])    # nth_element at size/2
        return out

    # --- states -------------------------------------------------------------
    def gt_state(self, stride, off_depth, nd):
        """Ground-truth poses with depth params = 1/(mean net field) as a rough GT."""
        x = np.zeros((self.N, stride))
        x[:, 0:3] = self.t; x[:, 3:6] = self.w_aa; x[:, 6] = self.phi
        for f in range(self.N):
            x[f, off_depth:off_depth + nd] = 1.0 / (np.exp(np.mean(np.log(self.field[f]))) * self.global_scale[f])
        return x

    def identity_state(self, stride, off_depth, nd, k=1):
        """The reference's starting point: position 0, identity orientation, default vFov
        (lib/DepthPhoto.h:23,43; lib/DepthPhoto.cpp:114-158), xform params 1 (spatial 0)."""
        x = np.zeros((self.N, stride))
        vfov = np.float32(0.666488587) if self.aspect > np.tan(0.508015513 / 2) / np.tan(0.666488587 / 2) else None
        if vfov is None:
            hf = np.tan(np.float32(0.508015513) / np.float32(2.0)); vf = 2 * np.arctan(hf / self.aspect)
        else:
            vf = float(vfov)
        x[:, 6] = np.tan(vf / 2.0)
        x[:, off_depth:off_depth + nd] = 1.0
        if k == 2:
            x[:, off_depth + 1:off_depth + nd:2] = 1.0   # GridDepthXform params_ all init 1.0 (:707)
        return x
