# first GPU shake-down: prints diagnostics instead of asserting
import numpy as np, time, sys, traceback
sys.path.insert(0, '.')
from robust_cvd_b200 import abi, solver, synthetic
from oracle import oracle
from tests import helpers
for name, ov in helpers.VARIANTS:
    try:
        sc, cfg, pairs, offs, rec, med = helpers.make_case(**ov)
        off_d, nd = helpers.layout_numbers(cfg)
        O = oracle.OracleProblem(cfg); G = solver.Problem(cfg)
        x = helpers.initial_state(sc, cfg, G.stride, off_d, nd)
        helpers.setup_problem(O, cfg, pairs, offs, rec, med, x); helpers.setup_problem(G, cfg, pairs, offs, rec, med, x)
        co, go = O.evaluate(True); cg, gg = G.evaluate(True)
        Ho = O.normal_matrix_dense(); Hg = G.normal_matrix_dense()
        print(name, "stride", G.stride, "cost", co, cg, "dgrad", np.abs(go-gg).max(), "dH", np.abs(Ho-Hg).max(), "Hmax", np.abs(Ho).max(), G.structure_info())
        U = Ho.shape[0]; rng = np.random.default_rng(3)
        S = 1/(1+np.sqrt(np.diag(Ho))); D2 = np.clip(S*S*np.diag(Ho),1e-6,1e32)/1e4; b = rng.normal(size=U)
        A = Ho*S[:,None]*S[None,:]+np.diag(D2); yr = np.linalg.solve(A,b)
        try:
            y = G.debug_linear_solve(S, D2, b)
            print("   linsolve res", np.linalg.norm(A@y-b)/np.linalg.norm(b), "err", np.linalg.norm(y-yr)/np.linalg.norm(yr))
        except Exception as e:
            print("   linsolve FAILED", e)
        opt = abi.default_solve_options(max_iterations=60)
        so = O.solve(opt); sg = G.solve(opt)
        xo, xg = O.get_state(), G.get_state()
        print("   solve oracle", so.termination, so.iterations, so.final_cost, so.message.decode(), "| gpu", sg.termination, sg.iterations, sg.final_cost, sg.message.decode(), "launches", sg.gpu_launches, "ms", round(sg.total_ms,2), "rel dx", np.linalg.norm(xo-xg)/np.linalg.norm(xo))
    except Exception as e:
        traceback.print_exc()
