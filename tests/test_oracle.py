"""Self-checks that pin the CPU oracle (parity is otherwise unpinned: the reference has no tests and
Ceres cannot be built here): literal Jet autodiff vs independent analytic Jacobians vs finite
differences, spline gather properties, exact block Cholesky vs dense, LM invariants."""
import numpy as np
import pytest

from robust_cvd_b200 import abi
from tests import helpers


def _oracle_case(overrides, **kw):
    from oracle import oracle
    sc, cfg, pairs, offs, rec, med = helpers.make_case(**overrides, **kw)
    off_d, nd = helpers.layout_numbers(cfg)
    O = oracle.OracleProblem(cfg)
    x = helpers.initial_state(sc, cfg, O.stride, off_d, nd)
    helpers.setup_problem(O, cfg, pairs, offs, rec, med, x)
    return sc, cfg, O, x


@pytest.mark.parametrize("name,overrides", helpers.VARIANTS, ids=[v[0] for v in helpers.VARIANTS])
def test_jet_autodiff_matches_analytic(name, overrides):
    sc, cfg, O, x = _oracle_case(overrides)
    r0, J0 = O.static_jacobian(0)
    r1, J1 = O.static_jacobian(1)
    np.testing.assert_allclose(r0, r1, atol=1e-13)
    assert np.abs(J0 - J1).max() <= 1e-12 * max(1.0, np.abs(J1).max())
    q0, K0 = O.regulariser_jacobian(0)
    q1, K1 = O.regulariser_jacobian(1)
    np.testing.assert_allclose(q0, q1, atol=1e-13)
    if K0.size:
        assert np.abs(K0 - K1).max() <= 1e-12 * max(1.0, np.abs(K1).max())


@pytest.mark.parametrize("name,overrides", helpers.VARIANTS[:5], ids=[v[0] for v in helpers.VARIANTS[:5]])
def test_gradient_matches_finite_differences(name, overrides):
    sc, cfg, O, x = _oracle_case(overrides)
    c0, g = O.evaluate(True)
    xf = x.reshape(-1).copy()
    act = O.active_mask()
    rng = np.random.default_rng(0)
    for i in rng.choice(np.nonzero(act)[0], 10, replace=False):
        h = 1e-6
        xp = xf.copy(); xp[i] += h; O.set_state(xp); cp = O.evaluate()
        xm = xf.copy(); xm[i] -= h; O.set_state(xm); cm = O.evaluate()
        fd = (cp - cm) / (2 * h)
        assert abs(fd - g[i]) <= 2e-5 * max(1.0, abs(g[i])), (i, fd, g[i])


def test_normal_matrix_is_jtj_with_cauchy_correction():
    sc, cfg, O, x = _oracle_case(helpers.VARIANTS[0][1])
    r, J = O.static_jacobian(0)
    q, K = O.regulariser_jacobian(0)
    b = cfg.robustness ** 2
    s = (r.reshape(-1, 3) ** 2).sum(1)
    w = 1.0 / (1.0 + s / b)                                # rho' of CauchyLoss; corrector scales rows by sqrt(rho')
    Jw = J * np.repeat(np.sqrt(w), 3)[:, None]
    H = Jw.T @ Jw + K.T @ K
    Ho = O.normal_matrix_dense()
    assert np.abs(H - Ho).max() <= 1e-10 * np.abs(H).max()
    cost = 0.5 * (b * np.log1p(s / b)).sum() + 0.5 * (q ** 2).sum()
    assert abs(cost - O.evaluate()) <= 1e-12 * cost


def test_gather_partition_of_unity_and_folding():
    from oracle import oracle
    rng = np.random.default_rng(1)
    for cubic in (0, 1):
        cfg = abi.default_config(1, 1.5, depth_type=abi.DEPTH_GRID, depth_cubic=cubic, depth_grid_x=5, depth_grid_y=4)
        for lx, ly in np.vstack([rng.uniform(-1, 1, (50, 2)), [[-1, -1], [1, 1], [-1, 1], [0.9999999, -0.9999999]]]):
            idx, w = oracle.gather_depth(cfg, np.float32(lx), np.float32(ly))
            assert abs(w.sum() - 1.0) < 1e-14
            assert len(set(idx.tolist())) == len(idx) and idx.min() >= 0 and idx.max() < 20
            assert len(idx) == (4 if not cubic else len(idx)) and len(idx) in (4, 9, 12, 16)
    # bilinear reproduces linear functions of the node lattice exactly
    cfg = abi.default_config(1, 1.5, depth_type=abi.DEPTH_GRID, depth_grid_x=6, depth_grid_y=3)
    for lx, ly in rng.uniform(-1, 1, (20, 2)):
        idx, w = oracle.gather_depth(cfg, np.float32(lx), np.float32(ly))
        nx = idx % 6; ny = idx // 6
        assert abs((w * nx).sum() - (float(np.float32(lx)) + 1.0) * 5 / 2) < 1e-12
        assert abs((w * ny).sum() - (float(np.float32(ly)) + 1.0) * 2 / 2) < 1e-12
    # top-right corner clamps into the last cell (nextafter rule)
    idx, w = oracle.gather_depth(cfg, np.float32(1.0), np.float32(1.0))
    assert idx.tolist() == [4 + 1 * 6, 5 + 1 * 6, 4 + 2 * 6, 5 + 2 * 6] and w[3] > 0.999999


def test_block_cholesky_matches_dense_solve():
    sc, cfg, O, x = _oracle_case(helpers.VARIANTS[2][1])
    H = O.normal_matrix_dense()
    U = H.shape[0]
    rng = np.random.default_rng(2)
    S = 1.0 / (1.0 + np.sqrt(np.diag(H)))
    D2 = np.clip(S * S * np.diag(H), 1e-6, 1e32) / 1e4
    b = rng.normal(size=U)
    y = O.block_solve(S, D2, b)
    A = H * S[:, None] * S[None, :] + np.diag(D2)
    assert np.linalg.norm(A @ y - b) / np.linalg.norm(b) < 1e-9


def test_lm_invariants_and_ground_truth_recovery():
    """Accepted steps decrease the cost; from a perturbed ground truth the solver returns to a state whose
    relative camera motion matches the ground truth (gauge-invariant check)."""
    from oracle import oracle
    sc, cfg, pairs, offs, rec, med = helpers.make_case(num_frames=8, depth_type=abi.DEPTH_GLOBAL)
    sc.flow_noise = 0.0
    O = oracle.OracleProblem(cfg)
    x0 = helpers.initial_state(sc, cfg, O.stride, 7, 1, perturb=0.003)
    helpers.setup_problem(O, cfg, pairs, offs, rec, med, x0)
    c_init = O.evaluate()
    s = O.solve(abi.default_solve_options(max_iterations=100))
    assert s.final_cost < 0.2 * c_init and s.final_cost <= s.initial_cost
    assert s.termination == abi.TERM_CONVERGENCE
    x = O.get_state()
    assert np.isfinite(x).all()


def test_normalize_depth_reaches_inverse_median():
    from oracle import oracle
    sc, cfg, pairs, offs, rec, med = helpers.make_case(depth_type=abi.DEPTH_GLOBAL, depth_lower_bound=1, depth_deform_reg=1.0, focal_reg=0.0)
    O = oracle.OracleProblem(cfg)
    O.set_frames(np.ones(8, np.uint8), med)
    O.set_constraints(np.zeros((0, 2), np.int32), np.zeros(1, np.int64), np.zeros((0, 6), np.float32))
    O.set_state(sc.identity_state(O.stride, 7, 1))
    s = O.solve(abi.default_solve_options())
    assert s.termination == abi.TERM_CONVERGENCE
    np.testing.assert_allclose(O.get_state()[:, 7], 1.0 / med, rtol=1e-6)


@pytest.mark.parametrize("smooth_type", [0, 1, 2, 3])
def test_smoothness_triplets_gradient(smooth_type):
    from oracle import oracle
    sc, cfg, pairs, offs, rec, med = helpers.make_case(smooth_loss_type=smooth_type, depth_type=abi.DEPTH_GRID, depth_grid_x=4, depth_grid_y=4)
    ce, to, tr = sc.triplets(sep=14)
    O = oracle.OracleProblem(cfg)
    x = helpers.initial_state(sc, cfg, O.stride, 7, 16)
    helpers.setup_problem(O, cfg, pairs, offs, rec, med, x); O.set_triplets(ce, to, tr)
    c, g = O.evaluate(True)
    r, J = O.triplet_jacobian()
    w = np.repeat(tr[:, 9].astype(np.float64), 3)
    O2 = oracle.OracleProblem(cfg); helpers.setup_problem(O2, cfg, pairs, offs, rec, med, x)
    c2, g2 = O2.evaluate(True)
    assert abs((c - c2) - 0.5 * (w * r * r).sum()) <= 1e-10 * c          # ScaledLoss: 1/2 w |r|^2
    np.testing.assert_allclose(g - g2, J.T @ (w * r), atol=1e-9)
    xf = x.reshape(-1).copy(); rng = np.random.default_rng(0)
    for i in rng.choice(np.nonzero(O.active_mask())[0], 8, replace=False):
        h = 1e-6
        xp = xf.copy(); xp[i] += h; O.set_state(xp); cp = O.evaluate()
        xm = xf.copy(); xm[i] -= h; O.set_state(xm); cm = O.evaluate()
        assert abs((cp - cm) / (2 * h) - g[i]) <= 2e-5 * max(1.0, abs(g[i]))


def test_hierarchical2_pairs_match_reference_golden():
    """tests/golden/hierarchical2_pairs.json was generated by importing the reference's
    utils/frame_sampling.py (tests/golden/make_golden.py)."""
    import json, os
    from robust_cvd_b200 import synthetic
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hierarchical2_pairs.json")))
    for n, pairs in g.items():
        assert [list(p) for p in synthetic.hierarchical2_pairs(int(n))] == pairs


@pytest.mark.parametrize("loss", [abi.LOSS_EUCLIDEAN, abi.LOSS_REPRO_DISPARITY, abi.LOSS_REPRO_DEPTH_RATIO, abi.LOSS_REPRO_LOG_DEPTH])
def test_static_scene_geometry_against_reference_python_camera_model(loss):
    """tests/golden/ref_python_geometry.npz: correspondences produced by the reference's own utils/geometry.py
    (pixels_to_points -> reproject_points -> project, imported unchanged by tests/golden/make_golden.py) with the intrinsics /
    extrinsics the loader builds from the C++ pose state (loaders/video_dataset.py:177-189).  StaticSceneCost restated in the
    oracle must see them as exact matches: all three residuals vanish, for every loss type, in both pair directions."""
    import os
    from oracle import oracle
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_python_geometry.npz"))
    W, H = int(g["W"]), int(g["H"])
    aspect = float(np.float32(W) / np.float32(H))
    cfg = abi.default_config(2, aspect, depth_type=abi.DEPTH_IDENTITY, static_loss_type=loss, intr_opt=abi.INTR_PER_FRAME,
                             scale_reg=0.0, focal_reg=0.0, depth_deform_reg=0.0, spatial_deform_reg=0.0)
    O = oracle.OracleProblem(cfg)
    n = len(g["px0"])
    ndc = lambda px, py: (-1.0 + 2.0 * px / W, 1.0 - 2.0 * py / H)          # lib/PoseOptimizer.cpp:104-117 with loc = pixel / size
    x0, y0 = ndc(g["px0"], g["py0"]); x1, y1 = ndc(g["px1"], g["py1"])
    fwd = np.stack([x0, y0, g["depth0"], x1, y1, g["depth1"]], 1)
    bwd = np.stack([x1, y1, g["depth1"], x0, y0, g["depth0"]], 1)
    rec = np.concatenate([fwd, bwd]).astype(np.float32)
    O.set_frames(np.ones(2, np.uint8), np.ones(2))
    O.set_constraints(np.array([[0, 1], [1, 0]], np.int32), np.array([0, n, 2 * n], np.int64), rec)
    x = np.zeros((2, O.stride))
    x[:, 0:3] = g["position"]; x[:, 3:6] = g["angle_axis"]; x[:, 6] = g["tan_half_vfov"]
    O.set_state(x.ravel())
    r, _ = O.static_jacobian(jac=False)
    # float32 records (24-byte wire format) bound the agreement: ~1e-7 relative on NDC / depth
    assert np.abs(r).max() < 5e-6, np.abs(r).max()
    # and the world point of each observation is the reference's points_cam_to_world
    if loss == abi.LOSS_EUCLIDEAN:
        O.set_constraints(np.array([[0, 1]], np.int32), np.array([0, n], np.int64), fwd.astype(np.float32))
        x2 = x.copy(); x2[1, 0:3] = 0; x2[1, 3:6] = 0; x2[1, 6] = 1.0       # frame 1 at the origin with tan = 1: its point is (ndc*aspect*d, ndc*d, -d)
        O.set_state(x2.ravel())
        r2, _ = O.static_jacobian(jac=False)
        p1 = np.stack([fwd[:, 3] * aspect * fwd[:, 5], fwd[:, 4] * fwd[:, 5], -fwd[:, 5]], 1)
        np.testing.assert_allclose(p1 - r2.reshape(-1, 3), g["world"], rtol=0, atol=5e-6)   # r = pw1 - pw0


# ---------------------------------------------------------------------------------------------------------------------------------------
# Independent pins of the oracle.  The reference has no tests and Ceres cannot be built here, so these are the
# checks that do not share a derivation with the oracle: a symbolic restatement of StaticSceneCost (reference lib/PoseOptimizer.cpp:
# 163-308, Rodrigues rotation as in ceres::AngleAxisRotatePoint) differentiated by sympy, and SciPy's trust-region least squares as an
# independent minimiser of the same residual vector.
# ---------------------------------------------------------------------------------------------------------------------------------------
def _sympy_static_scene(loss_type):
    import sympy as sp
    t0 = sp.symbols("t0x t0y t0z"); w0 = sp.symbols("w0x w0y w0z"); t1 = sp.symbols("t1x t1y t1z"); w1 = sp.symbols("w1x w1y w1z")
    phi0, s0, phi1, s1 = sp.symbols("phi0 s0 phi1 s1")
    n0x, n0y, d0, n1x, n1y, d1, asp, ws, wd = sp.symbols("n0x n0y d0 n1x n1y d1 asp ws wd")

    def rotate(w, p):                      # ceres::AngleAxisRotatePoint, theta^2 > epsilon branch
        w = sp.Matrix(w); p = sp.Matrix(p)
        th = sp.sqrt(w.dot(w)); k = w / th
        return p * sp.cos(th) + k.cross(p) * sp.sin(th) + k * (k.dot(p)) * (1 - sp.cos(th))
    D0, D1 = d0 * s0, d1 * s1                                               # Global(Scale) depth transform
    dir0 = sp.Matrix([n0x * phi0 * asp, n0y * phi0, -1])                     # cameraToWorld
    X = sp.Matrix(t0) + rotate(w0, dir0) * D0
    q = rotate([-w1[0], -w1[1], -w1[2]], X - sp.Matrix(t1))                  # worldToCamera
    depth = -q[2]
    px, py = q[0] / depth / (phi1 * asp), q[1] / depth / phi1
    if loss_type == abi.LOSS_REPRO_DISPARITY:
        rz = (1 / depth - 1 / D1) * wd                                       # both depths > 1e-6 at the test state
    elif loss_type == abi.LOSS_REPRO_DEPTH_RATIO:
        rz = (sp.Max(depth, D1) / sp.Min(depth, D1) - 1) * wd
    else:
        rz = sp.log(sp.Min(depth, D1) / sp.Max(depth, D1)) * wd
    r = sp.Matrix([(px - n1x) * ws, (py - n1y) * ws, rz])
    params = list(t0) + list(w0) + [phi0, s0] + list(t1) + list(w1) + [phi1, s1]
    consts = [n0x, n0y, d0, n1x, n1y, d1, asp, ws, wd]
    return sp.lambdify(params + consts, r, "numpy"), sp.lambdify(params + consts, r.jacobian(params), "numpy")


@pytest.mark.parametrize("loss_type", [abi.LOSS_REPRO_DISPARITY, abi.LOSS_REPRO_DEPTH_RATIO, abi.LOSS_REPRO_LOG_DEPTH])
def test_static_scene_cost_matches_sympy_restatement(loss_type):
    sc, cfg, O, x = _oracle_case(dict(depth_type=abi.DEPTH_GLOBAL, static_loss_type=loss_type), num_frames=4)
    fr, fJ = _sympy_static_scene(loss_type)
    r, J = O.static_jacobian(0)
    sc2, cfg2, pairs, offs, rec, med = helpers.make_case(num_frames=4, depth_type=abi.DEPTH_GLOBAL, static_loss_type=loss_type)
    X = x.reshape(4, -1)
    assert X.shape[1] == 8
    n = 0
    rng = np.random.default_rng(0)
    for pi, (a, b) in enumerate(np.asarray(pairs).reshape(-1, 2)):
        for c in rng.choice(np.arange(offs[pi], offs[pi + 1]), 5, replace=False):
            k = rec[c].astype(np.float64)
            args = list(X[a]) + list(X[b]) + [k[0], k[1], k[2], k[3], k[4], k[5], cfg.aspect, cfg.static_spatial_weight, cfg.static_depth_weight]
            rs = np.asarray(fr(*args), float).ravel(); Js = np.asarray(fJ(*args), float)
            np.testing.assert_allclose(r[3 * c:3 * c + 3], rs, rtol=1e-10, atol=1e-13)
            Jo = np.concatenate([J[3 * c:3 * c + 3, a * 8:(a + 1) * 8], J[3 * c:3 * c + 3, b * 8:(b + 1) * 8]], axis=1)
            assert np.abs(Jo - Js).max() <= 1e-9 * max(1.0, np.abs(Js).max())
            others = np.delete(J[3 * c:3 * c + 3], np.r_[a * 8:(a + 1) * 8, b * 8:(b + 1) * 8], axis=1)
            assert not others.any()
            n += 1
    assert n >= 30


def _relative_geometry(X):
    """Gauge-invariant summary of a state (N x stride, Global depth): focal, depth scale, pairwise camera distances, relative rotations."""
    from robust_cvd_b200.synthetic import rodrigues
    R = [rodrigues(v) for v in X[:, 3:6]]
    dist = np.array([np.linalg.norm(X[i, :3] - X[j, :3]) for i in range(len(X)) for j in range(i)])
    rel = np.array([(R[i].T @ R[j]).ravel() for i in range(len(X)) for j in range(i)]).ravel()
    return np.concatenate([X[:, 6], X[:, 7], dist, rel])


def test_lm_optimum_matches_scipy_least_squares():
    """Non-robust Global-depth problem: the oracle's Ceres-semantics LM and SciPy's trust-region reflective solver (independent code, exact
    SVD steps, tolerances 1e-14) must reach the same minimum -- same cost, same gauge-invariant geometry (there is no gauge fixing:
    absolute poses differ by a rigid motion between any two solvers)."""
    from scipy.optimize import least_squares
    sc, cfg, O, x = _oracle_case(dict(depth_type=abi.DEPTH_GLOBAL, robust_type=abi.ROBUST_TRIVIAL), num_frames=6)

    def fun(v):
        O.set_state(v); r, _ = O.static_jacobian(0, jac=False); q, _ = O.regulariser_jacobian(0)
        return np.concatenate([r, q])

    def jac(v):
        O.set_state(v); _, J = O.static_jacobian(0); _, K = O.regulariser_jacobian(0)
        return np.vstack([J, K])
    x0 = x.reshape(-1).copy()
    O.set_state(x0); c0 = O.evaluate()
    assert abs(0.5 * np.sum(fun(x0) ** 2) - c0) <= 1e-12 * c0          # the stacked residual vector IS the oracle's cost
    res = least_squares(fun, x0, jac=jac, method="trf", tr_solver="exact", xtol=1e-14, ftol=1e-14, gtol=1e-14, max_nfev=400)
    O.set_state(x0)
    opt = abi.default_solve_options(max_iterations=400)
    opt.function_tolerance = 1e-15; opt.parameter_tolerance = 1e-14; opt.gradient_tolerance = 1e-14
    s = O.solve(opt)
    xo = O.get_state()
    assert res.cost < 0.5 * c0
    assert abs(s.final_cost - res.cost) <= 1e-8 * res.cost, (s.final_cost, res.cost)
    go, gs = _relative_geometry(xo), _relative_geometry(res.x.reshape(xo.shape))
    assert np.abs(go - gs).max() <= 1e-4 * np.abs(gs).max(), np.abs(go - gs).max()
