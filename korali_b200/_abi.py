"""ctypes mirror of include/kcma.h (the C ABI of libkcma.so).

The binder is generic over the symbol prefix so that the CPU oracle (oracle/libokcma.so, prefix
``okcma_``; test infrastructure only) can be driven through the same vocabulary by the tests.
Nothing in this module loads the oracle.
"""
import ctypes as C
import math
import numpy as np

KCMA_ABI_VERSION = 1
MU_TYPES = {"Linear": 0, "Equal": 1, "Logarithmic": 2, "Proportional": 3}
OBJECTIVES = {"NegSphere": 0, "NegRosenbrock": 1, "NegAckley": 2, "NegEllipsoid": 3, "NegSumSq": 4,
              "NegSphereSin2": 5, "External": 100}
CONSTRAINTS = {"None": 0, "HalfSpace": 1, "External": 100}
INJ_Z, INJ_BDZ, INJ_X, INJ_F, INJ_BD, INJ_GRAD = 0, 1, 2, 3, 4, 5

_dp = C.POINTER(C.c_double)


class KcmaCfg(C.Structure):
    """struct kcma_cfg (include/kcma.h)."""
    _fields_ = [
        ("abi_version", C.c_uint32), ("reserved0", C.c_uint32),
        ("n", C.c_uint64), ("population_size", C.c_uint64), ("mu_value", C.c_uint64),
        ("mu_type", C.c_int32), ("diagonal_covariance", C.c_int32), ("mirrored_sampling", C.c_int32),
        ("is_sigma_bounded", C.c_int32),
        ("initial_sigma_cumulation_factor", C.c_double), ("initial_damp_factor", C.c_double),
        ("initial_cumulative_covariance", C.c_double),
        ("viability_population_size", C.c_uint64), ("viability_mu_value", C.c_uint64),
        ("max_covariance_matrix_corrections", C.c_uint64),
        ("target_success_rate", C.c_double), ("covariance_matrix_adaption_strength", C.c_double),
        ("normal_vector_learning_rate", C.c_double), ("global_success_learning_rate", C.c_double),
        ("max_infeasible_resamplings", C.c_uint64), ("seed", C.c_uint64),
        ("objective", C.c_int32), ("constraint_family", C.c_int32), ("n_constraints", C.c_uint64),
        ("objective_coef", _dp), ("constraint_shift", _dp),
        ("lower_bound", _dp), ("upper_bound", _dp), ("initial_value", _dp), ("initial_stddev", _dp),
        ("min_stddev_update", _dp),
        ("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32), ("keep_population", C.c_int32),
        ("use_gradient_information", C.c_int32), ("reserved1", C.c_int32), ("gradient_step_size", C.c_double),
        ("granularity", _dp),
    ]


class KcmaError(RuntimeError):
    """Raised for every non-zero return code; carries the KORALI_LOG_ERROR-style message."""


def _as_dp(a):
    return a.ctypes.data_as(_dp)


class Handle:
    """A solver handle of libkcma.so (prefix 'kcma_') — or of the oracle (prefix 'okcma_') in tests."""

    def __init__(self, lib, prefix, **kw):
        self._lib, self._p = lib, prefix
        self._keep = []
        cfg = KcmaCfg()
        self._fn("cfg_defaults", None, [C.POINTER(KcmaCfg)])(C.byref(cfg))
        n = int(kw["n"])
        for k, v in kw.items():
            if k in ("lower_bound", "upper_bound", "initial_value", "initial_stddev", "min_stddev_update",
                     "objective_coef", "constraint_shift", "granularity"):
                if v is None:
                    continue
                want = int(kw.get("n_constraints", 0)) if k == "constraint_shift" else n
                arr = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (want,)))
                self._keep.append(arr)
                setattr(cfg, k, _as_dp(arr))
            elif k == "mu_type":
                cfg.mu_type = MU_TYPES[v] if isinstance(v, str) else int(v)
            elif k == "objective":
                cfg.objective = OBJECTIVES[v] if isinstance(v, str) else int(v)
            elif k == "constraint_family":
                cfg.constraint_family = CONSTRAINTS[v] if isinstance(v, str) else int(v)
            else:
                setattr(cfg, k, v)
        self.n = n
        self._h = C.c_void_p()
        create = self._fn("create", C.c_int, [C.POINTER(KcmaCfg), C.POINTER(C.c_void_p)])
        if create(C.byref(cfg), C.byref(self._h)) != 0:
            raise KcmaError(self._fn("last_error", C.c_char_p, [C.c_void_p])(None).decode())

    def _fn(self, name, restype, argtypes):
        f = getattr(self._lib, self._p + name)
        f.restype, f.argtypes = restype, argtypes
        return f

    def _check(self, rc):
        if rc != 0:
            raise KcmaError(self._fn("last_error", C.c_char_p, [C.c_void_p])(self._live()).decode())

    def _live(self):
        """The handle, or an error instead of a NULL pointer into the library once close() was called."""
        if not getattr(self, "_h", None):
            raise KcmaError("the solver handle is closed")
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            self._fn("destroy", None, [C.c_void_p])(self._h)
            self._h = None

    __del__ = close

    # --- generation loop
    def ask(self): self._check(self._fn("ask", C.c_int, [C.c_void_p])(self._live()))
    def eval(self): self._check(self._fn("eval", C.c_int, [C.c_void_p])(self._live()))
    def tell(self): self._check(self._fn("tell", C.c_int, [C.c_void_p])(self._live()))
    def run_generation(self): self._check(self._fn("run_generation", C.c_int, [C.c_void_p])(self._live()))

    def run(self, max_generations):
        done = C.c_uint64(0)
        self._check(self._fn("run", C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)])(
            self._h, int(max_generations), C.byref(done)))
        return done.value

    def check_termination(self):
        fin, reason = C.c_int(0), C.c_char_p()
        self._check(self._fn("check_termination", C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_char_p)])(
            self._h, C.byref(fin), C.byref(reason)))
        return bool(fin.value), (reason.value or b"").decode()

    def take_warnings(self):
        return self._fn("take_warnings", C.c_char_p, [C.c_void_p])(self._live()).decode()

    def inject(self, kind, data):
        a = np.ascontiguousarray(data, dtype=np.float64).ravel()
        self._check(self._fn("inject", C.c_int, [C.c_void_p, C.c_int, _dp, C.c_size_t])(self._live(), kind, _as_dp(a), a.size))

    # --- state
    def get(self, key):
        """Array by Korali key (1-D float64)."""
        f = self._fn("get_array", C.c_int, [C.c_void_p, C.c_char_p, _dp, C.c_size_t, C.POINTER(C.c_size_t)])
        cnt = C.c_size_t(0)
        self._check(f(self._live(), key.encode(), None, 0, C.byref(cnt)))
        out = np.empty(cnt.value, dtype=np.float64)
        self._check(f(self._live(), key.encode(), _as_dp(out), out.size, C.byref(cnt)))
        return out

    def set(self, key, data):
        a = np.ascontiguousarray(data, dtype=np.float64).ravel()
        self._check(self._fn("set_array", C.c_int, [C.c_void_p, C.c_char_p, _dp, C.c_size_t])(self._live(), key.encode(), _as_dp(a), a.size))

    def get_index(self, key):
        f = self._fn("get_index_array", C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint64), C.c_size_t, C.POINTER(C.c_size_t)])
        cnt = C.c_size_t(0)
        self._check(f(self._live(), key.encode(), None, 0, C.byref(cnt)))
        out = np.empty(cnt.value, dtype=np.uint64)
        self._check(f(self._live(), key.encode(), out.ctypes.data_as(C.POINTER(C.c_uint64)), out.size, C.byref(cnt)))
        return out

    def scalar(self, key):
        v = C.c_double(math.nan)
        self._check(self._fn("get_scalar", C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_double)])(self._live(), key.encode(), C.byref(v)))
        return v.value

    def set_scalar(self, key, value):
        self._check(self._fn("set_scalar", C.c_int, [C.c_void_p, C.c_char_p, C.c_double])(self._live(), key.encode(), float(value)))
