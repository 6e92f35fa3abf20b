"""Generates the committed golden fixtures.  Run in the build container (needs /root/reference for the
pure-Python pieces of the reference; the solver goldens come from the CPU oracle because the reference's
C++ path cannot be built here -- see DESIGN.md "Oracle")."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def golden_pairs():
    sys.path.insert(0, "/root/reference")
    from utils import frame_sampling as fs
    from utils.frame_range import FrameRange, OptionalSet
    out = {}
    for n in (2, 3, 8, 17, 64, 300):
        fr = FrameRange(frame_range=OptionalSet(), num_frames=n)
        pairs = fs.SamplePairs.sample([fs.SamplePairsOptions(mode=fs.SamplePairsMode.HIERARCHICAL2)], frame_range=fr, two_way=True)
        out[str(n)] = [[int(p[0]), int(p[1])] for p in pairs]
    json.dump(out, open(os.path.join(HERE, "hierarchical2_pairs.json"), "w"))


def golden_solver():
    from robust_cvd_b200 import abi
    from oracle import oracle
    from tests import helpers
    out = {}
    for name in ("bilinear_perframe_disp", "bicubic_shared_ratio_bicubicwarp", "global_fixed_log_bilinearwarp"):
        ov = dict(helpers.VARIANTS)[name]
        sc, cfg, pairs, offs, rec, med = helpers.make_case(**ov)
        off_d, nd = helpers.layout_numbers(cfg)
        O = oracle.OracleProblem(cfg)
        x = helpers.initial_state(sc, cfg, O.stride, off_d, nd)
        helpers.setup_problem(O, cfg, pairs, offs, rec, med, x)
        cost, g = O.evaluate(True)
        Hd = np.diag(O.normal_matrix_dense()).copy()
        s = O.solve(abi.default_solve_options(max_iterations=60))
        out[name + "/x0"] = x; out[name + "/cost"] = np.array(cost); out[name + "/grad"] = g; out[name + "/hdiag"] = Hd
        out[name + "/final_cost"] = np.array(s.final_cost); out[name + "/iterations"] = np.array(s.iterations); out[name + "/x_final"] = O.get_state()
    np.savez_compressed(os.path.join(HERE, "oracle_solver_golden.npz"), **out)


if __name__ == "__main__":
    golden_pairs()
    golden_solver()
    print("golden fixtures written")
