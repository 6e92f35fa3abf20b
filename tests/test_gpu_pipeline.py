"""GPU tests of the drop-in boundary: the optimize_poses() call sequence of the reference's
pose_optimization.py:177-240 through lib_python, each optimisation step replayed on the CPU oracle
from the identical problem arrays; committed golden vectors; dense transform kernels."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robust_cvd_b200", "host"))

from robust_cvd_b200 import abi, synthetic, synthetic_files  # noqa: E402
from tests import helpers  # noqa: E402

pytestmark = pytest.mark.gpu
CV_32FC3 = 21


def _oracle_replay(d, max_iterations):
    from oracle import oracle
    cfg = abi.Config.from_buffer_copy(d["config"])
    O = oracle.OracleProblem(cfg)
    O.set_frames(d["in_range"], d["median"], d["adaptive"] if d["adaptive"].size else None)
    O.set_constraints(d["pair_frames"].reshape(-1, 2), d["offsets"], d["records"].reshape(-1, 6))
    if d["trip_centers"].size:
        O.set_triplets(d["trip_centers"], d["trip_offsets"], d["trip_records"].reshape(-1, 10))
    O.set_state(d["state"])
    s = O.solve(abi.default_solve_options(max_iterations=max_iterations))
    return O.get_state(), s


def test_optimize_poses_sequence_matches_oracle_step_by_step(tmp_path):
    import lib_python as lp
    root = str(tmp_path / "scene")
    sc = synthetic.Scene(8, 128, 96, seed=3)
    synthetic_files.write_scene(sc, root)
    v = lp.DepthVideo(); lp.DepthVideoImporter.importVideo(v, root, False)
    v.createColorStream("full", "color_full", ".png", CV_32FC3); v.createColorStream("down", "color_down", ".raw", CV_32FC3)
    v.createDepthStream("depth_midas2", "depth_midas2", [-1, -1])
    fp = lp.FlowConstraintsParams(); fp.frameRange.resolve(v.numFrames(), True)
    fc = lp.FlowConstraintsCollection(v, fp); fc.setStaticFlagFromDynamicMask(8)
    proc = lp.DepthVideoProcessor(v)
    params = lp.DepthVideoProcessor.Params()
    params.depthStream = v.numDepthStreams() - 1
    params.frameRange.fromString("0-7"); params.poseOptimizer.frameRange.fromString("0-7")
    params.poseOptimizer.maxIterations = 40
    params.op = lp.DepthVideoProcessor.Op.ResetDepthXforms
    params.depthXformDesc.type = lp.XformType.Depth; params.depthXformDesc.depthType = lp.DepthXformType.Global; params.depthXformDesc.valueXform = lp.ValueXformType.Scale
    proc.process(params)
    params.op = lp.DepthVideoProcessor.Op.ResetSpatialXforms
    params.spatialXformDesc.type = lp.XformType.Spatial; params.spatialXformDesc.spatialType = lp.SpatialXformType.Identity; params.spatialXformDesc.valueXform = lp.ValueXformType.Scale
    proc.process(params)
    ds = v.depthStream(params.depthStream)
    # --- normalizeDepth ---
    opt = lp.DepthVideoPoseOptimizer(v, params.depthStream)
    d = opt._buildProblem(params.poseOptimizer, fc, 0.0, True)
    xo, so = _oracle_replay(d, 40)
    proc.normalizeDepth(params, fc)
    s0 = xo.reshape(8, -1)[0, 7]
    for f in range(8):      # first frame's transform copied to all (lib/PoseOptimizer.cpp:1127-1138)
        assert abs(ds.frame(f).depthXform().params()[0] - s0) <= 1e-6 * abs(s0)
    # --- coarse-to-fine steps, one at a time (numSteps = 1 per call), replayed on the oracle ---
    grids = [(1, 1), (6, 4), (12, 7), (17, 10)]       # ctfLong 17 / ctfShort 10, landscape (lib/PoseOptimizer.cpp:795-802, :858-863)
    p1 = lp.DepthVideoProcessor.Params(); p1.depthStream = params.depthStream
    p1.poseOptimizer = params.poseOptimizer; p1.poseOptimizer.numSteps = 1; p1.poseOptimizer.coarseToFine = False
    for step, (gx, gy) in enumerate(grids):
        if step > 0:
            sp = lp.DepthVideoProcessor.Params(); sp.depthStream = params.depthStream
            sp.depthXformDesc.parse(f"Grid(Scale, Linear, {gx}, {gy}, 1)"); proc.gridXformSplit(sp)
        opt = lp.DepthVideoPoseOptimizer(v, params.depthStream)
        d = opt._buildProblem(p1.poseOptimizer, fc, p1.poseOptimizer.depthDeformRegFinal, False)
        xo, so = _oracle_replay(d, 40)
        proc.optimizePoses(p1, fc)
        opt2 = lp.DepthVideoPoseOptimizer(v, params.depthStream)
        xg = opt2._buildProblem(p1.poseOptimizer, fc, 0.1, False)["state"].reshape(8, -1)
        xo = xo.reshape(8, -1)
        # depth-transform params are carried in double: direct comparison; poses went through the float32 write-back
        nd = gx * gy if step > 0 else 1
        np.testing.assert_allclose(xg[:, 7:7 + nd], xo[:, 7:7 + nd], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(xg[:, :3], xo[:, :3].astype(np.float32), rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(xg[:, 6], xo[:, 6], rtol=1e-4)
    assert ds.depthXformDesc().str() == "Grid(Scale, Linear, 17, 10, 1)"
    # full poseOptimization() in one call from the same start gives the same schedule
    v.save()
    assert os.path.getsize(f"{root}/video.dat") > 8 * (170 * 8)
    # results as the training loop reads them (loaders/video_dataset.py:153-217)
    f = ds.frame(3)
    pm = f.depthXform().paramMap(f); wp = f.spatialXform().warp(ds.height(), ds.width())
    assert pm.shape == (96, 128) and pm.dtype == np.float64 and wp.shape == (96, 128, 2) and not wp.any()
    R = np.stack([f.extrinsics.right(), f.extrinsics.up(), f.extrinsics.backward()], 1)
    np.testing.assert_allclose(R.T @ R, np.eye(3), atol=1e-5)


def test_smoothness_loss_through_lib_python(tmp_path):
    """smoothStaticWeight / smoothDynamicWeight > 0 enables the scene-flow smoothness triplets (lib/PoseOptimizer.cpp:899-901)."""
    import lib_python as lp
    root = str(tmp_path / "scene")
    sc = synthetic.Scene(8, 128, 96, seed=5)
    synthetic_files.write_scene(sc, root)
    v = lp.DepthVideo(); lp.DepthVideoImporter.importVideo(v, root, False)
    v.createColorStream("down", "color_down", ".raw", CV_32FC3); v.createDepthStream("depth_midas2", "depth_midas2", [-1, -1])
    fp = lp.FlowConstraintsParams(); fp.frameRange.resolve(v.numFrames(), True)
    fc = lp.FlowConstraintsCollection(v, fp); fc.setStaticFlagFromDynamicMask(8)
    proc = lp.DepthVideoProcessor(v)
    params = lp.DepthVideoProcessor.Params(); params.depthStream = 0
    params.poseOptimizer.frameRange.fromString("0-7"); params.poseOptimizer.maxIterations = 40
    params.poseOptimizer.numSteps = 1; params.poseOptimizer.coarseToFine = False
    params.poseOptimizer.smoothStaticWeight = 0.5; params.poseOptimizer.smoothDynamicWeight = 0.25
    params.depthXformDesc.parse("Grid(Scale, Linear, 5, 4, 1)"); proc.resetDepthXforms(params)
    opt = lp.DepthVideoPoseOptimizer(v, 0)
    d = opt._buildProblem(params.poseOptimizer, fc, 0.1, False)
    assert d["trip_centers"].tolist() == [1, 2, 3, 4, 5, 6] and d["trip_records"].size > 600
    assert np.all(d["trip_records"].reshape(-1, 10)[:, 9] == np.float32(0.5))       # no dynamic masks -> all static
    xo, so = _oracle_replay(d, 40)
    proc.optimizePoses(params, fc)
    xg = lp.DepthVideoPoseOptimizer(v, 0)._buildProblem(params.poseOptimizer, fc, 0.1, False)["state"].reshape(8, -1)
    np.testing.assert_allclose(xg[:, 7:27], xo.reshape(8, -1)[:, 7:27], rtol=1e-4, atol=1e-7)


def test_device_memory_is_returned_after_processor_calls():
    """rcvd_trim_device_memory: the stream-ordered pool keeps freed blocks cached across solver calls; the host side hands them
    back after every DepthVideoProcessor operation because the reference's fine-tuning stage (PyTorch) shares the GPU."""
    import torch
    from robust_cvd_b200 import solver
    sc, cfg, pairs, offs, rec, med = helpers.make_case(num_frames=16, depth_type=abi.DEPTH_GRID, depth_grid_x=16, depth_grid_y=12)
    off_d, nd = helpers.layout_numbers(cfg)
    solver.lib().rcvd_trim_device_memory(0)
    free0 = torch.cuda.mem_get_info(0)[0]
    G = solver.Problem(cfg)
    helpers.setup_problem(G, cfg, pairs, offs, rec, med, helpers.initial_state(sc, cfg, G.stride, off_d, nd))
    G.solve(abi.default_solve_options(max_iterations=3))
    G.close()
    cached = free0 - torch.cuda.mem_get_info(0)[0]
    assert cached > 32 << 20                                  # the pool kept the factor storage (~100 MB at npad 208)
    assert solver.lib().rcvd_trim_device_memory(0) == 0
    assert free0 - torch.cuda.mem_get_info(0)[0] < cached // 4


def test_full_pose_optimization_call(tmp_path):
    """DepthVideoProcessor.optimizePoses with the default 4-step coarse-to-fine schedule in one call."""
    import lib_python as lp
    root = str(tmp_path / "scene")
    sc = synthetic.Scene(8, 128, 96, seed=4)
    synthetic_files.write_scene(sc, root)
    v = lp.DepthVideo(); lp.DepthVideoImporter.importVideo(v, root, False)
    v.createColorStream("down", "color_down", ".raw", CV_32FC3)
    v.createDepthStream("depth_midas2", "depth_midas2", [-1, -1])
    fp = lp.FlowConstraintsParams(); fp.frameRange.resolve(v.numFrames(), True)
    fc = lp.FlowConstraintsCollection(v, fp); fc.setStaticFlagFromDynamicMask(8)
    proc = lp.DepthVideoProcessor(v)
    params = lp.DepthVideoProcessor.Params(); params.depthStream = 0
    params.poseOptimizer.frameRange.fromString("0-7"); params.poseOptimizer.maxIterations = 50
    params.depthXformDesc.parse("Global(Scale)"); proc.resetDepthXforms(params)
    proc.normalizeDepth(params, fc)
    proc.optimizePoses(params, fc)
    ds = v.depthStream(0)
    assert ds.depthXformDesc().str() == "Grid(Scale, Linear, 17, 10, 1)"
    pos = np.stack([ds.frame(i).extrinsics.position for i in range(8)])
    assert np.isfinite(pos).all() and np.abs(pos).max() > 0      # cameras moved away from the identity start
    scales = np.array(ds.frame(5).depthXform().params())
    assert scales.shape == (170,) and (scales > 0).all()


def test_cuda_path_against_committed_golden_vectors():
    from robust_cvd_b200 import solver
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_solver_golden.npz"))
    for name in ("bilinear_perframe_disp", "bicubic_shared_ratio_bicubicwarp", "global_fixed_log_bilinearwarp"):
        ov = dict(helpers.VARIANTS)[name]
        sc, cfg, pairs, offs, rec, med = helpers.make_case(**ov)
        G = solver.Problem(cfg)
        helpers.setup_problem(G, cfg, pairs, offs, rec, med, g[name + "/x0"])
        c, gr = G.evaluate(True)
        assert abs(c - float(g[name + "/cost"])) <= 1e-11 * abs(c)
        assert np.abs(gr - g[name + "/grad"]).max() <= 1e-9 * max(1.0, np.abs(gr).max())
        assert np.abs(np.diag(G.normal_matrix_dense()) - g[name + "/hdiag"]).max() <= 1e-9 * g[name + "/hdiag"].max()
        s = G.solve(abi.default_solve_options(max_iterations=60))
        assert abs(s.final_cost - float(g[name + "/final_cost"])) <= 1e-6 * s.final_cost
        assert abs(s.iterations - int(g[name + "/iterations"])) <= 2
        assert np.linalg.norm(G.get_state() - g[name + "/x_final"]) <= 1e-4 * np.linalg.norm(g[name + "/x_final"])


def test_dense_transform_kernels():
    """rcvd_depth_apply / rcvd_depth_param_map / rcvd_spatial_warp against the oracle's gathers at the
    dense pixel->NDC map x = -1 + x*2/(w-1), y = 1 - y*2/(h-1) (lib/DepthMapTransform.cpp:397-407)."""
    from oracle import oracle
    from robust_cvd_b200 import solver
    rng = np.random.default_rng(0)
    h, w = 24, 40
    src = rng.uniform(0.5, 3.0, (h, w)).astype(np.float32)
    for cubic, k in ((0, 1), (1, 1), (1, 2)):
        cfg = abi.default_config(1, 1.5, depth_type=abi.DEPTH_GRID, depth_cubic=cubic, depth_grid_x=5, depth_grid_y=4,
                                 value_xform=abi.VALUE_SCALESHIFT if k == 2 else abi.VALUE_SCALE)
        p = rng.uniform(0.5, 1.5, 20 * k)
        out = solver.depth_apply(cfg, p, src); pm = solver.depth_param_map(cfg, p, h, w)
        xs = np.float32(2.0) / (np.float32(w) - np.float32(1.0)); ys = np.float32(2.0) / (np.float32(h) - np.float32(1.0))
        for (y, x) in [(0, 0), (h - 1, w - 1), (5, 17), (h - 1, 0), (11, 39)]:
            lx = np.float32(-1.0) + np.float32(x) * xs; ly = np.float32(1.0) - np.float32(y) * ys
            idx, wt = oracle.gather_depth(cfg, lx, ly)
            if k == 1:
                ref = float(sum((float(src[y, x]) * p[i]) * wi for i, wi in zip(idx, wt))); refp = float(sum(p[i] * wi for i, wi in zip(idx, wt)))
                assert abs(pm[y, x] - refp) < 1e-13
            else:
                ref = float(sum((float(src[y, x]) * p[2 * i] + p[2 * i + 1]) * wi for i, wi in zip(idx, wt)))
                assert abs(pm[y, x, 0] - sum(p[2 * i] * wi for i, wi in zip(idx, wt))) < 1e-13
            assert abs(out[y, x] - np.float32(ref)) <= 1e-6 * abs(ref)
    cfg = abi.default_config(1, 1.5, spatial_type=abi.SPATIAL_BICUBIC_GRID, spatial_grid_x=4, spatial_grid_y=3)
    sp = rng.normal(0, 0.01, 24)
    wp = solver.spatial_warp(cfg, sp, h, w)
    lx = np.float32(-1.0) + np.float32(7) * (np.float32(2.0) / (np.float32(w) - np.float32(1.0))); ly = np.float32(1.0) - np.float32(3) * (np.float32(2.0) / (np.float32(h) - np.float32(1.0)))
    idx, wt = oracle.gather_spatial(cfg, lx, ly)
    np.testing.assert_allclose(wp[3, 7], [sum(sp[2 * i] * wi for i, wi in zip(idx, wt)), sum(sp[2 * i + 1] * wi for i, wi in zip(idx, wt))], rtol=1e-6, atol=1e-9)


def test_filter_depth_sequence_through_lib_python(tmp_path):
    """The call sequence of the reference's filter_depth (pose_optimization.py:292-326): create stream, Op.Copy,
    Op.FlowGuidedFilter, saveDepth, save -- against the float32 restatement of the reference loop; then video.dat is re-read."""
    import lib_python as lp
    from oracle import host_ref
    root = str(tmp_path / "scene")
    N, W, H = 6, 48, 32
    sc = synthetic.Scene(N, W, H, seed=5, motion=0.05, rot_deg=0.5)
    synthetic_files.write_scene(sc, root)
    v = lp.DepthVideo(); lp.DepthVideoImporter.importVideo(v, root, False)
    v.createColorStream("down", "color_down", ".raw", CV_32FC3)
    v.createDepthStream("depth_midas2", "depth_midas2", [-1, -1])
    src_id = v.numDepthStreams() - 1
    src = v.depthStream(src_id)
    # give the frames distinct cameras and a non-trivial depth transform
    from scipy.spatial.transform import Rotation
    for f in range(N):
        df = src.frame(f)
        e = df.extrinsics; e.position = np.asarray(sc.t[f], np.float32)
        q = Rotation.from_matrix(sc.R[f]).as_quat().astype(np.float32)
        e.orientation = lp._makeQuat(*[float(c) for c in q])
        df.extrinsics = e
    proc = lp.DepthVideoProcessor(v)
    rp = lp.DepthVideoProcessor.Params(); rp.depthStream = src_id
    rp.depthXformDesc.type = lp.XformType.Depth; rp.depthXformDesc.depthType = lp.DepthXformType.Global; rp.depthXformDesc.valueXform = lp.ValueXformType.Scale
    proc.resetDepthXforms(rp)
    for f in range(N):
        src.frame(f).depthXform().params()[0] = 1.0 / float(sc.global_scale[f])
        src.frame(f).clearXformedCache()
    # --- filter_depth(radius = 2) ---
    dst_id = v.numDepthStreams()
    v.createDepthStream(src.name() + "_filtered", "depth_midas2_filtered", [src.width(), src.height()])
    params = lp.DepthVideoProcessor.Params()
    assert (params.spatialRadius, params.frameRadius, params.median, params.farConnections) == (0, 2, False, False)   # lib/Processor.h:66-72
    params.frameRange.fromString("1-5")
    params.op = lp.DepthVideoProcessor.Op.Copy; params.sourceDepthStream = src_id; params.depthStream = dst_id
    proc.process(params)
    np.testing.assert_array_equal(v.depthStream(dst_id).frame(3).depth(), src.frame(3).depth())
    params.op = lp.DepthVideoProcessor.Op.FlowGuidedFilter; params.frameRadius = 2
    proc.process(params)
    got = np.stack([np.array(v.depthStream(dst_id).frame(f).depth()) for f in range(1, 6)])
    # reference loop on the same inputs (frames 0..5 are all inside the windows: base = max(0, 1 - 2) = 0)
    depth = np.stack([np.array(src.frame(f).depth()) for f in range(N)])
    cams = np.zeros((N, 9), np.float32)
    for f in range(N):
        df = src.frame(f); e = df.extrinsics
        cams[f, :3] = e.position; cams[f, 3:7] = [e.orientation.x(), e.orientation.y(), e.orientation.z(), e.orientation.w()]
        cams[f, 7] = df.intrinsics.hFov; cams[f, 8] = df.intrinsics.vFov
    fwd = np.zeros((N, H, W, 2), np.float32); fwm = np.zeros((N, H, W), np.uint8); bwd = np.zeros_like(fwd); bwm = np.zeros_like(fwm)
    import cv2
    for f in range(N - 1):
        fwd[f] = synthetic_files.read_raw(os.path.join(root, "flow", f"flow_{f:06d}_{f + 1:06d}.raw")); fwm[f] = cv2.imread(os.path.join(root, "flow_mask", f"mask_{f:06d}_{f + 1:06d}.png"), cv2.IMREAD_GRAYSCALE)
        bwd[f + 1] = synthetic_files.read_raw(os.path.join(root, "flow", f"flow_{f + 1:06d}_{f:06d}.raw")); bwm[f + 1] = cv2.imread(os.path.join(root, "flow_mask", f"mask_{f + 1:06d}_{f:06d}.png"), cv2.IMREAD_GRAYSCALE)
    want = host_ref.flow_guided_filter(depth, cams, fwd, fwm, bwd, bwm, first_out=1, num_out=5, frame_radius=2, spatial_radius=0, median=False, inv_aspect=v.invAspect())
    np.testing.assert_allclose(got, want, rtol=1e-5)
    assert np.abs(got - depth[1:6]).max() > 1e-4
    # --- saveDepth + save, then re-read both ---
    v.saveDepth(dst_id); v.save()
    disp = synthetic_files.read_raw(os.path.join(root, "depth_midas2_filtered", "depth", "frame_000003.raw"))
    np.testing.assert_allclose(disp, (np.float32(1.0) / got[2]).astype(np.float32), rtol=1e-6)
    v2 = lp.DepthVideo(); v2.load(root)
    assert v2.numFrames() == N and v2.numDepthStreams() == 2 and v2.numColorStreams() == 1 and v2.width() == v.width() and abs(v2.invAspect() - v.invAspect()) == 0
    assert v2.depthStream(0).depthXformDesc().str() == src.depthXformDesc().str() and v2.depthStream(1).name() == "depth_midas2_filtered"
    for f in range(N):
        a, b = v2.depthStream(0).frame(f), src.frame(f)
        np.testing.assert_array_equal(np.asarray(a.extrinsics.position), np.asarray(b.extrinsics.position))
        assert list(a.depthXform().params()) == list(b.depthXform().params())
        assert a.intrinsics.vFov == b.intrinsics.vFov
    np.testing.assert_allclose(np.array(v2.depthStream(1).frame(3).depth()), got[2], rtol=2e-6)   # disparity round trip


def test_filter_far_connections_median_and_partial_range(tmp_path):
    """Op.FlowGuidedFilter with farConnections + median + spatialRadius 1 on the frame range 2-4 of a 7-frame video:
    the host side keeps every frame on the device (far targets may lie anywhere) and windows reach back before the range."""
    import cv2
    import lib_python as lp
    from oracle import host_ref
    from scipy.spatial.transform import Rotation
    root = str(tmp_path / "scene")
    N, W, H = 7, 40, 28
    sc = synthetic.Scene(N, W, H, seed=11, motion=0.05, rot_deg=0.5)
    synthetic_files.write_scene(sc, root)       # hierarchical2 pairs: consecutive frames both ways + (0,2),(2,0),(0,4),(4,0),(2,4),(4,2),(4,6),(6,4)
    v = lp.DepthVideo(); lp.DepthVideoImporter.importVideo(v, root, False)
    v.createColorStream("down", "color_down", ".raw", CV_32FC3)
    v.createDepthStream("depth_midas2", "depth_midas2", [-1, -1])
    src = v.depthStream(0)
    for f in range(N):
        df = src.frame(f); e = df.extrinsics; e.position = np.asarray(sc.t[f], np.float32)
        e.orientation = lp._makeQuat(*[float(c) for c in Rotation.from_matrix(sc.R[f]).as_quat().astype(np.float32)]); df.extrinsics = e
    v.createDepthStream("filtered", "depth_filtered", [src.width(), src.height()])
    proc = lp.DepthVideoProcessor(v)
    params = lp.DepthVideoProcessor.Params(); params.frameRange.fromString("2-4")
    params.op = lp.DepthVideoProcessor.Op.Copy; params.sourceDepthStream = 0; params.depthStream = 1; proc.process(params)
    params.op = lp.DepthVideoProcessor.Op.FlowGuidedFilter; params.frameRadius = 1; params.spatialRadius = 1; params.median = True; params.farConnections = True
    proc.process(params)
    got = np.stack([np.array(v.depthStream(1).frame(f).depth()) for f in (2, 3, 4)])
    depth = np.stack([np.array(src.frame(f).depth()) for f in range(N)])
    cams = np.zeros((N, 9), np.float32)
    for f in range(N):
        df = src.frame(f); e = df.extrinsics
        cams[f, :3] = e.position; cams[f, 3:7] = [e.orientation.x(), e.orientation.y(), e.orientation.z(), e.orientation.w()]; cams[f, 7] = df.intrinsics.hFov; cams[f, 8] = df.intrinsics.vFov

    def fm(a, b):
        return (synthetic_files.read_raw(os.path.join(root, "flow", f"flow_{a:06d}_{b:06d}.raw")), cv2.imread(os.path.join(root, "flow_mask", f"mask_{a:06d}_{b:06d}.png"), cv2.IMREAD_GRAYSCALE))
    fwd = np.zeros((N, H, W, 2), np.float32); fwm = np.zeros((N, H, W), np.uint8); bwd = np.zeros_like(fwd); bwm = np.zeros_like(fwm)
    for f in range(N - 1):
        fwd[f], fwm[f] = fm(f, f + 1); bwd[f + 1], bwm[f + 1] = fm(f + 1, f)
    # far connections of frames 2..4 (frameRadius 1, range last = 4): flow files (frame, fi) with fi outside [max(0, frame-1), min(4, frame+1)], sorted
    far = []
    for name in sorted(os.listdir(os.path.join(root, "flow"))):
        a, b = int(name[5:11]), int(name[12:18])
        if 2 <= a <= 4 and (b < max(0, a - 1) or b > min(4, a + 1)):
            far.append((a, b))
    assert (2, 0) in far and (4, 6) in far and (4, 5) in far          # a target beyond the range end counts as "far" (f1 is clamped to the range)
    ff = np.stack([fm(a, b)[0] for a, b in far]); fmk = np.stack([fm(a, b)[1] for a, b in far])
    want = host_ref.flow_guided_filter(depth, cams, fwd, fwm, bwd, bwm, first_out=2, num_out=3, frame_radius=1, spatial_radius=1, median=True,
                                       inv_aspect=v.invAspect(), far_pairs=far, far_flow=ff, far_mask=fmk)
    assert np.isclose(got, want, rtol=1e-5).mean() >= 0.995, np.abs(got - want).max()
    np.testing.assert_array_equal(np.array(v.depthStream(1).frame(1).depth() is None), True)     # frames outside the range were not created


def test_constraint_builder_on_a_frame_sub_range(tmp_path, monkeypatch):
    """FlowConstraintsParams.frameRange = 2-5 of 8 frames: local frame indices of the GPU builder have an offset; same lists as the host builder."""
    import lib_python as lp
    root = str(tmp_path / "scene")
    sc = synthetic.Scene(8, 64, 48, seed=13)
    synthetic_files.write_scene(sc, root)

    def run(which):
        monkeypatch.setenv("RCVD_CONSTRAINT_BUILDER", which)
        v = lp.DepthVideo(); lp.DepthVideoImporter.importVideo(v, root, False)
        v.createColorStream("down", "color_down", ".raw", CV_32FC3); v.createDepthStream("depth_midas2", "depth_midas2", [-1, -1])
        fp = lp.FlowConstraintsParams(); fp.frameRange.fromString("2-5"); fp.frameRange.resolve(v.numFrames(), True); fp.doNotUseCache = True
        fc = lp.FlowConstraintsCollection(v, fp)
        return {k: np.asarray(a[0]) for k, a in fc._pairs().items()}, {k: np.asarray(a[0]) for k, a in fc._triplets().items()}
    gp, gt = run("gpu"); hp, ht = run("host")
    assert sorted(gp) == sorted(hp) and all(2 <= a <= 5 and 2 <= b <= 5 for a, b in gp) and len(gp) >= 6 and sorted(gt) == [3, 4]
    for k in gp:
        np.testing.assert_array_equal(gp[k], hp[k])
    for k in gt:
        np.testing.assert_array_equal(gt[k], ht[k])
