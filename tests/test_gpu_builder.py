"""GPU flow-constraint builder: rcvd_build_constraints vs the numpy / cv2 restatement of the reference's
sequential sampler (oracle/host_ref.py, lib/FlowConstraints.cpp:352-465) and vs the sequential host builder through
lib_python -- bit-exact constraint lists (integer / index work)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robust_cvd_b200", "host"))

from robust_cvd_b200 import solver, synthetic, synthetic_files  # noqa: E402

pytestmark = pytest.mark.gpu
CV_32FC3, CV_8UC1 = 21, 0


def _scene(tmp_path, N=6, W=96, H=64, seed=7, dynamic=True):
    root = str(tmp_path / "scene")
    sc = synthetic.Scene(N, W, H, seed=seed, motion=0.04, rot_deg=0.5)
    masks = None
    if dynamic:
        rng = np.random.default_rng(seed)
        masks = []
        for i in range(N):
            m = np.full((H, W), 255, np.uint8); cx, cy = rng.integers(20, W - 20), rng.integers(15, H - 15); m[cy - 7:cy + 7, cx - 9:cx + 9] = 0; masks.append(m)
    pairs = synthetic_files.write_scene(sc, root, dynamic_masks=masks)
    return sc, root, pairs


@pytest.mark.parametrize("sep", [10, 3, 1, 0])
def test_builder_matches_numpy_cv2_restatement(tmp_path, sep):
    from oracle import host_ref
    sc, root, pairs = _scene(tmp_path, dynamic=False)
    N, W, H = sc.N, sc.w, sc.h
    color = np.stack([synthetic_files.read_raw(os.path.join(root, "color_down", f"frame_{i:06d}.raw")) for i in range(N)])
    import cv2
    sel = [p for p in pairs][:10]
    flow = np.stack([synthetic_files.read_raw(os.path.join(root, "flow", f"flow_{a:06d}_{b:06d}.raw")) for a, b in sel])
    mask = np.stack([cv2.imread(os.path.join(root, "flow_mask", f"mask_{a:06d}_{b:06d}.png"), cv2.IMREAD_GRAYSCALE) for a, b in sel])
    l0 = solver.lib().rcvd_builder_launch_count()
    poff, pc, toff, tc = solver.build_constraints(color, sel, flow, mask, sep, float(sc.inv_aspect32))
    assert solver.lib().rcvd_builder_launch_count() > l0
    assert poff[0] == 0 and poff[-1] == len(pc) and len(tc) == 0
    import lib_python as lp
    for k, (a, b) in enumerate(sel):
        # priorities of the restatement come from the real cv2.cornerMinEigenVal at every separation (0 = every candidate survives,
        # 1 / 3 keep thousands of near-equal priorities): the CUDA corner score is bit-equal to OpenCV's
        want, _ = host_ref.pair_constraints(color[a], flow[k], mask[k], sep, sc.inv_aspect32)
        got = pc[poff[k]:poff[k + 1]]
        assert got.shape == want.shape, (k, got.shape, want.shape)
        np.testing.assert_array_equal(got, want)
    if sep == 10:
        assert 20 < (poff[1] - poff[0]) < 200          # the disc sampler really thinned the candidates
        assert solver.lib().rcvd_builder_last_rounds() >= 2


def _lists(fc):
    return {k: np.asarray(v[0]) for k, v in fc._pairs().items()}, {k: np.asarray(v[0]) for k, v in fc._triplets().items()}


@pytest.mark.parametrize("sep,min_dyn", [(10, -1.0), (4, 3.0)])
def test_gpu_builder_equals_host_builder_through_lib_python(tmp_path, monkeypatch, sep, min_dyn):
    import lib_python as lp
    sc, root, pairs = _scene(tmp_path, dynamic=True)

    def run(which):
        monkeypatch.setenv("RCVD_CONSTRAINT_BUILDER", which)
        v = lp.DepthVideo(); lp.DepthVideoImporter.importVideo(v, root, False)
        v.createColorStream("down", "color_down", ".raw", CV_32FC3); v.createColorStream("dynamic_mask", "dynamic_mask", ".png", CV_8UC1)
        v.createDepthStream("depth_midas2", "depth_midas2", [-1, -1])
        fp = lp.FlowConstraintsParams(); fp.frameRange.resolve(v.numFrames(), True); fp.matchSeparation = sep; fp.minDynamicDistance = min_dyn; fp.doNotUseCache = True
        return _lists(lp.FlowConstraintsCollection(v, fp))
    l0 = solver.lib().rcvd_builder_launch_count()
    gp, gt = run("gpu")
    assert solver.lib().rcvd_builder_launch_count() > l0          # the CUDA path really ran
    l1 = solver.lib().rcvd_builder_launch_count()
    hp, ht = run("host")
    assert solver.lib().rcvd_builder_launch_count() == l1
    assert gp.keys() == hp.keys() and gt.keys() == ht.keys() and len(gp) == len(pairs) and len(gt) == sc.N - 2
    assert sum(len(v) for v in gp.values()) > 100 and sum(len(v) for v in gt.values()) > 20
    for k in gp:
        np.testing.assert_array_equal(gp[k], hp[k], err_msg=f"pair {k}")
    for k in gt:
        np.testing.assert_array_equal(gt[k], ht[k], err_msg=f"triplet {k}")


def test_device_distance_transform_and_static_flags_match_host(tmp_path, monkeypatch):
    """A2 on the device (rcvd_static_flags): the 5x5 fixed-point chamfer distance of every dynamic mask is bit-equal to the host restatement
    of cv::distanceTransform(DIST_L2, 5) (and within the fixed-point/IPP float difference of the real cv2), and the static flags of all pair
    and triplet constraints equal the sequential host pass through lib_python (reference lib/FlowConstraints.cpp:573-660)."""
    import cv2
    import lib_python as lp
    sc, root, pairs = _scene(tmp_path, dynamic=True)
    N, W, H = sc.N, sc.w, sc.h
    masks = np.stack([cv2.imread(os.path.join(root, "dynamic_mask", f"frame_{i:06d}.png"), cv2.IMREAD_GRAYSCALE) for i in range(N)])
    rng = np.random.default_rng(9)
    stress = (rng.uniform(size=(3, 200, 333)) > 0.02).astype(np.uint8) * 255            # scattered single dynamic pixels, width not a multiple of the scan chunk
    wide = np.full((1, 40, 700), 255, np.uint8); wide[0, 20, 5] = 0; wide[0, 3, 690] = 0  # distances that travel through several 256-column scan chunks
    for mm in (masks, stress, wide):
        l0 = solver.lib().rcvd_static_flag_launch_count()
        _, _, dist = solver.static_flags(mm, 8.0, want_distance=True)
        assert solver.lib().rcvd_static_flag_launch_count() > l0
        for f in range(mm.shape[0]):
            b = np.where(mm[f] < 127, 0, 255).astype(np.uint8)
            np.testing.assert_array_equal(dist[f], np.asarray(lp._distanceTransformL2_5(b)))
            np.testing.assert_allclose(dist[f], cv2.distanceTransform(b, cv2.DIST_L2, 5), rtol=1e-5, atol=1e-3)

    def run(which):
        monkeypatch.setenv("RCVD_CONSTRAINT_BUILDER", which)
        v = lp.DepthVideo(); lp.DepthVideoImporter.importVideo(v, root, False)
        v.createColorStream("down", "color_down", ".raw", CV_32FC3); v.createColorStream("dynamic_mask", "dynamic_mask", ".png", CV_8UC1)
        v.createDepthStream("depth_midas2", "depth_midas2", [-1, -1])
        fp = lp.FlowConstraintsParams(); fp.frameRange.resolve(v.numFrames(), True); fp.doNotUseCache = True
        fc = lp.FlowConstraintsCollection(v, fp)
        fc.setStaticFlagFromDynamicMask(8)
        return {k: np.asarray(v_[1]) for k, v_ in fc._pairs().items()}, {k: np.asarray(v_[1]) for k, v_ in fc._triplets().items()}
    l0 = solver.lib().rcvd_static_flag_launch_count()
    gp, gt = run("gpu")
    assert solver.lib().rcvd_static_flag_launch_count() > l0
    l1 = solver.lib().rcvd_static_flag_launch_count()
    hp, ht = run("host")
    assert solver.lib().rcvd_static_flag_launch_count() == l1
    n_dyn = 0
    for k in gp:
        np.testing.assert_array_equal(gp[k], hp[k], err_msg=f"pair {k}"); n_dyn += int((~gp[k].astype(bool)).sum())
    for k in gt:
        np.testing.assert_array_equal(gt[k], ht[k], err_msg=f"triplet {k}")
    assert n_dyn > 0 and len(gp) == len(pairs)


def test_builder_against_committed_golden():
    """Disc sampler goldens (tests/golden/host_restatement_golden.npz): the stored priorities come from cv2, whose last bits
    differ from the CUDA operator, so only the separation-4 list (no near-ties among its ~40 survivors) is compared exactly."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "host_restatement_golden.npz"))
    poff, pc, _, _ = solver.build_constraints(g["sampler_color"][None], [(0, 0)], g["sampler_flow"][None], g["sampler_mask"][None], 4, 0.75)
    np.testing.assert_array_equal(pc, g["sampler_sep4"])


def test_builder_empty_and_bad_arguments():
    color = np.zeros((2, 8, 8, 3), np.float32)
    poff, pc, toff, tc = solver.build_constraints(color, np.zeros((0, 2), np.int32), None, None, 10, 1.0)
    assert len(pc) == 0 and len(tc) == 0
    with pytest.raises(RuntimeError):
        solver.build_constraints(color, [(0, 5)], np.zeros((1, 8, 8, 2), np.float32), np.full((1, 8, 8), 255, np.uint8), 10, 1.0)
    # all-masked pair: zero constraints, valid offsets
    poff, pc, _, _ = solver.build_constraints(color, [(0, 1)], np.zeros((1, 8, 8, 2), np.float32), np.zeros((1, 8, 8), np.uint8), 3, 1.0)
    assert list(poff) == [0, 0]
