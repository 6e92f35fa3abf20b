import sys; sys.path.insert(0,'.')
import numpy as np
from robust_cvd_b200 import abi, solver
from oracle import oracle
from tests import helpers
ov = dict(helpers.VARIANTS)["global_euclid"]
sc, cfg, pairs, offs, rec, med = helpers.make_case(**ov)
off_d, nd = helpers.layout_numbers(cfg)
O = oracle.OracleProblem(cfg); G = solver.Problem(cfg)
x = helpers.initial_state(sc, cfg, G.stride, off_d, nd)
helpers.setup_problem(O, cfg, pairs, offs, rec, med, x); helpers.setup_problem(G, cfg, pairs, offs, rec, med, x)
opt = abi.default_solve_options(max_iterations=60, verbose=1)
so = O.solve(opt); sg = G.solve(opt)
print(so.final_cost, sg.final_cost, so.iterations, sg.iterations, so.message, sg.message)
