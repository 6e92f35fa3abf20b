"""Writes a synthetic.Scene as a robust_cvd working directory (SURVEY.md sections 7.0 / 8f-3):
frames.txt, color_down/frame_%06d.raw, depth_<tag>/depth/frame_%06d.raw (disparity),
flow/flow_%06d_%06d.raw, flow_mask/mask_%06d_%06d.png, flow_list.json -- the on-disk formats of
lib/core/CvUtil.cpp:25-42 (.raw), lib/Importer.cpp:197-238 (frames.txt), flow.py:53-74 (flow_list.json).
"""
import json
import os
import struct

import numpy as np

CV_8UC1, CV_32FC1, CV_32FC2, CV_32FC3 = 0, 5, 13, 21


def write_raw(path, arr):
    arr = np.ascontiguousarray(arr)
    cn = 1 if arr.ndim == 2 else arr.shape[2]
    depth = {np.dtype(np.uint8): 0, np.dtype(np.int32): 4, np.dtype(np.float32): 5, np.dtype(np.float64): 6}[arr.dtype]
    with open(path, "wb") as f:
        f.write(struct.pack("<iiiQ", arr.shape[0], arr.shape[1], depth + ((cn - 1) << 3), arr.dtype.itemsize * cn))
        f.write(arr.tobytes())


def read_raw(path):
    with open(path, "rb") as f:
        rows, cols, typ, es = struct.unpack("<iiiQ", f.read(20))
        dt = {0: np.uint8, 4: np.int32, 5: np.float32, 6: np.float64}[typ & 7]
        cn = (typ >> 3) + 1
        a = np.frombuffer(f.read(), dtype=dt).reshape(rows, cols, cn)
    return a[:, :, 0] if cn == 1 else a


def write_png_gray(path, img):
    """Minimal 8-bit grayscale PNG writer (no cv2 dependency)."""
    import zlib
    img = np.ascontiguousarray(img, np.uint8); h, w = img.shape
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

    def chunk(t, d):
        c = struct.pack(">I", len(d)) + t + d
        return c + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def texture(scene, f):
    """Seeded, corner-rich BGR float image for frame f (values in [0,1])."""
    rng = np.random.default_rng(scene.seed * 1000 + f)
    base = rng.uniform(0, 1, (scene.h // 4 + 2, scene.w // 4 + 2, 3)).astype(np.float32)
    img = np.kron(base, np.ones((4, 4, 1), np.float32))[:scene.h, :scene.w]
    img = 0.7 * img + 0.3 * rng.uniform(0, 1, img.shape).astype(np.float32)
    return np.ascontiguousarray(img.astype(np.float32))


def _write_pairs(scene, root, pairs, seeds):
    """Flow + mask files of a list of pairs; returns their flow_list rows.  seeds: one rng seed per pair, or a shared Generator."""
    iy, ix = np.mgrid[0:scene.h, 0:scene.w]
    rows = []
    for k, (a, b) in enumerate(pairs):
        rng = seeds if isinstance(seeds, np.random.Generator) else np.random.default_rng(seeds[k])
        fx1, fy1, ok = scene.flow(a, b, ix.ravel(), iy.ravel(), rng)
        flow = np.stack([fx1 - ix.ravel().astype(np.float32), fy1 - iy.ravel().astype(np.float32)], axis=-1).reshape(scene.h, scene.w, 2).astype(np.float32)
        inside = ok & (fx1 >= 0) & (fx1 <= scene.w - 1) & (fy1 >= 0) & (fy1 <= scene.h - 1)
        mask = (inside.reshape(scene.h, scene.w) * 255).astype(np.uint8)
        write_raw(os.path.join(root, "flow", f"flow_{a:06d}_{b:06d}.raw"), flow)
        write_png_gray(os.path.join(root, "flow_mask", f"mask_{a:06d}_{b:06d}.png"), mask)
        rows.append([int(a), int(b), float(mask.mean() / 255.0)])
    return rows


def _write_frames(scene, root, depth_tag, frames, dynamic_masks):
    for i in frames:
        write_raw(os.path.join(root, "color_down", f"frame_{i:06d}.raw"), texture(scene, i))
        depth = scene.depth_image(i)
        write_raw(os.path.join(root, depth_tag, "depth", f"frame_{i:06d}.raw"), (np.float32(1.0) / depth).astype(np.float32))
        if dynamic_masks is not None:
            write_png_gray(os.path.join(root, "dynamic_mask", f"frame_{i:06d}.png"), dynamic_masks[i])


def write_scene(scene, root, depth_tag="depth_midas2", pairs=None, full_size=None, dynamic_masks=None, workers=1):
    """workers > 1: frames and pairs are written by a process pool (flow noise is then seeded per pair instead of drawn from one
    sequential generator -- a different but equally deterministic realisation)."""
    from .synthetic import hierarchical2_pairs
    os.makedirs(root, exist_ok=True)
    for d in ("color_down", "color_full", f"{depth_tag}/depth", "flow", "flow_mask"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    if dynamic_masks is not None:
        os.makedirs(os.path.join(root, "dynamic_mask"), exist_ok=True)
    W, H = full_size or (scene.w, scene.h)
    with open(os.path.join(root, "frames.txt"), "w") as f:
        f.write(f"{scene.N}\n{W}\n{H}\n" + "".join(f"{i / 30.0:.6f}\n" for i in range(scene.N)))
    if pairs is None:
        pairs = hierarchical2_pairs(scene.N)
    pairs = list(pairs)
    if workers <= 1:
        _write_frames(scene, root, depth_tag, range(scene.N), dynamic_masks)
        rows = _write_pairs(scene, root, pairs, np.random.default_rng(scene.seed + 777))
    else:
        import multiprocessing as mp
        ctx = mp.get_context("fork")
        fchunks = [list(range(scene.N))[i::workers] for i in range(workers)]
        pidx = [list(range(len(pairs)))[i::workers] for i in range(workers)]
        with ctx.Pool(workers) as pool:
            jobs = [pool.apply_async(_write_frames, (scene, root, depth_tag, fc, dynamic_masks)) for fc in fchunks if fc]
            pjobs = [pool.apply_async(_write_pairs, (scene, root, [pairs[k] for k in ix], [scene.seed * 100003 + 777 + k for k in ix])) for ix in pidx if ix]
            for j in jobs:
                j.get()
            got = {}
            for ix, j in zip([ix for ix in pidx if ix], pjobs):
                for k, r in zip(ix, j.get()):
                    got[k] = r
        rows = [got[k] for k in range(len(pairs))]
    with open(os.path.join(root, "flow_list.json"), "w") as f:
        json.dump([["first", "second", "ratio"]] + rows, f)
    return pairs
