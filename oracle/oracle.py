"""ctypes wrapper of the CPU oracle (oracle/librcvd_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never imported by the product
package robust_cvd_b200/.
"""
import ctypes as C
import os
import subprocess
import numpy as np

from robust_cvd_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "librcvd_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_last_error.restype = C.c_char_p
        L.orc_problem_create.argtypes = [C.POINTER(abi.Config), C.POINTER(C.c_void_p)]
        L.orc_problem_destroy.argtypes = [C.c_void_p]
        L.orc_frame_stride.argtypes = [C.POINTER(abi.Config)]
        _LIB = L
        L.orc_set_threads(C.c_int32(effective_cpus()))
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class OracleProblem:
    """Array-level problem: same calls as robust_cvd_b200.solver.Problem."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.L = lib()
        self.h = C.c_void_p()
        rc = self.L.orc_problem_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise RuntimeError(self.L.orc_last_error().decode())
        self.N = cfg.num_frames
        self.stride = self.L.orc_frame_stride(C.byref(cfg))
        self.U = self.N * self.stride
        self.num_constraints = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_problem_destroy(self.h)
            self.h = None

    def set_frames(self, in_range=None, median_depth=None, adaptive_weights=None):
        ir = None if in_range is None else np.ascontiguousarray(in_range, np.uint8)
        md = None if median_depth is None else np.ascontiguousarray(median_depth, np.float64)
        aw = None if adaptive_weights is None else np.ascontiguousarray(adaptive_weights, np.float64)
        self.L.orc_problem_set_frames(self.h, _p(ir, C.c_uint8), _p(md, C.c_double), _p(aw, C.c_double))

    def set_constraints(self, pair_frames, offsets, records):
        pf = np.ascontiguousarray(pair_frames, np.int32).reshape(-1, 2)
        off = np.ascontiguousarray(offsets, np.int64)
        rec = np.ascontiguousarray(records, np.float32).reshape(-1, 6)
        assert off.shape[0] == pf.shape[0] + 1 and off[-1] == rec.shape[0]
        self.num_constraints = int(rec.shape[0])
        self.L.orc_problem_set_constraints(self.h, C.c_int32(pf.shape[0]), _p(pf, C.c_int32), _p(off, C.c_int64), _p(rec, C.c_float))

    def set_triplets(self, centers, offsets, records):
        """Scene-flow smoothness constraints: centers[T] (middle frame), offsets[T+1], records[n][10] =
        3 x (ndc.x, ndc.y, depth) + weight (smoothStaticWeight or smoothDynamicWeight)."""
        ce = np.ascontiguousarray(centers, np.int32); off = np.ascontiguousarray(offsets, np.int64)
        rec = np.ascontiguousarray(records, np.float32).reshape(-1, 10)
        assert off.shape[0] == ce.shape[0] + 1 and off[-1] == rec.shape[0]
        self.num_triplets = int(rec.shape[0])
        self.L.orc_problem_set_triplets(self.h, C.c_int32(ce.shape[0]), _p(ce, C.c_int32), _p(off, C.c_int64), _p(rec, C.c_float))

    def triplet_jacobian(self):
        r = np.zeros(3 * self.num_triplets, np.float64); J = np.zeros((3 * self.num_triplets, self.U), np.float64)
        self.L.orc_triplet_jacobian(self.h, _p(r, C.c_double), _p(J, C.c_double))
        return r, J

    def set_state(self, x):
        x = np.ascontiguousarray(x, np.float64).reshape(-1)
        assert x.size == self.U
        self.L.orc_problem_set_state(self.h, _p(x, C.c_double))

    def get_state(self):
        x = np.empty(self.U, np.float64)
        self.L.orc_problem_get_state(self.h, _p(x, C.c_double))
        return x.reshape(self.N, self.stride)

    def set_jacobian_mode(self, mode):
        """0: analytic, 1: Jet<4> passes (mirrors DynamicAutoDiffCostFunction<.,4>)."""
        self.L.orc_set_jacobian_mode(self.h, C.c_int32(mode))

    def evaluate(self, gradient=False):
        cost = C.c_double()
        g = np.zeros(self.U, np.float64) if gradient else None
        self.L.orc_evaluate(self.h, C.byref(cost), _p(g, C.c_double))
        return (cost.value, g) if gradient else cost.value

    def normal_matrix_dense(self):
        H = np.zeros((self.U, self.U), np.float64)
        self.L.orc_normal_matrix_dense(self.h, _p(H, C.c_double))
        return H

    def static_jacobian(self, mode=0, jac=True):
        r = np.zeros(3 * self.num_constraints, np.float64)
        J = np.zeros((3 * self.num_constraints, self.U), np.float64) if jac else None
        self.L.orc_static_jacobian(self.h, C.c_int32(mode), _p(r, C.c_double), _p(J, C.c_double))
        return r, J

    def regulariser_jacobian(self, mode=0):
        n = self.L.orc_regulariser_jacobian(self.h, C.c_int32(mode), None, None, C.c_int32(0))
        r = np.zeros(n, np.float64)
        J = np.zeros((n, self.U), np.float64)
        self.L.orc_regulariser_jacobian(self.h, C.c_int32(mode), _p(r, C.c_double), _p(J, C.c_double), C.c_int32(n))
        return r, J

    def active_mask(self):
        m = np.zeros(self.U, np.uint8)
        self.L.orc_active_mask(self.h, _p(m, C.c_uint8))
        return m.astype(bool)

    def solve(self, options=None):
        opt = options or abi.default_solve_options()
        s = abi.SolveSummary()
        rc = self.L.orc_solve(self.h, C.byref(opt), C.byref(s))
        if rc != 0:
            raise RuntimeError("oracle solve failed")
        return s

    def time_iteration(self, radius=1e4):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self.L.orc_time_iteration(self.h, C.c_double(radius), C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def block_solve(self, S, D2, b):
        S = np.ascontiguousarray(S, np.float64); D2 = np.ascontiguousarray(D2, np.float64)
        b = np.ascontiguousarray(b, np.float64); y = np.zeros_like(b)
        rc = self.L.orc_block_solve(self.h, _p(S, C.c_double), _p(D2, C.c_double), _p(b, C.c_double), _p(y, C.c_double))
        if rc != 0:
            raise RuntimeError("block solve failed")
        return y


def gather_depth(cfg, lx, ly):
    idx = np.zeros(16, np.int32); w = np.zeros(16, np.float64)
    n = lib().orc_gather_depth(C.byref(cfg), C.c_float(lx), C.c_float(ly), _p(idx, C.c_int32), _p(w, C.c_double))
    return idx[:n].copy(), w[:n].copy()


def gather_spatial(cfg, lx, ly):
    idx = np.zeros(16, np.int32); w = np.zeros(16, np.float64)
    n = lib().orc_gather_spatial(C.byref(cfg), C.c_float(lx), C.c_float(ly), _p(idx, C.c_int32), _p(w, C.c_double))
    return idx[:n].copy(), w[:n].copy()


def effective_cpus():
    """CPUs this container may actually use: min(visible CPUs, cgroup quota)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, q // per))
            break
        except Exception:
            continue
    return n


def set_threads(n):
    lib().orc_set_threads(C.c_int32(n))


def get_threads():
    return lib().orc_get_threads()
