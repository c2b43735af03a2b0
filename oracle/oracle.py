"""ctypes loader of the CPU oracle (oracle/libokcma.so). TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess
import numpy as np
from korali_b200._abi import Handle, _as_dp, _dp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libokcma.so")
        if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("okcma.c", "odea.c", "omocma.c")):
            build()
        _LIB = C.CDLL(path)
    return _LIB


OBJ_CB = C.CFUNCTYPE(None, C.c_void_p, _dp, C.c_uint64, _dp)
CON_CB = C.CFUNCTYPE(None, C.c_void_p, _dp, C.c_uint64, _dp, C.c_uint64)


class Oracle(Handle):
    def __init__(self, **kw):
        kw.pop("device", None)
        super().__init__(lib(), "okcma_", **kw)

    def set_objective(self, fn):
        """fn(x: ndarray[N]) -> float, called once per sample like the reference's Conduit."""
        def tramp(_u, x, n, out):
            out[0] = float(fn(np.ctypeslib.as_array(x, shape=(n,)).copy()))
        self._obj_cb = OBJ_CB(tramp)
        self._fn("set_objective_callback", None, [C.c_void_p, OBJ_CB, C.c_void_p])(self._h, self._obj_cb, None)

    def set_constraints(self, fns):
        def tramp(_u, x, n, out, nc):
            xv = np.ctypeslib.as_array(x, shape=(n,)).copy()
            for c in range(nc):
                out[c] = float(fns[c](xv))
        self._con_cb = CON_CB(tramp)
        self._fn("set_constraints_callback", None, [C.c_void_p, CON_CB, C.c_void_p])(self._h, self._con_cb, None)


class OracleDEA:
    """The DEA oracle (oracle/odea.c) behind the vocabulary of korali_b200._dea.DeaHandle."""

    def __new__(cls, **kw):
        from korali_b200._dea import DeaHandle
        kw.pop("device", None)
        h = DeaHandle(lib(), "odea_", **kw)

        def set_objective(fn, h=h):
            def tramp(_u, x, n, out):
                out[0] = float(fn(np.ctypeslib.as_array(x, shape=(n,)).copy()))
            h._obj_cb = OBJ_CB(tramp)
            h._fn("set_objective_callback", None, [C.c_void_p, OBJ_CB, C.c_void_p])(h._h, h._obj_cb, None)
        h.set_objective = set_objective
        return h


def sort_index(f):
    f = np.ascontiguousarray(f, dtype=np.float64)
    out = np.empty(f.size, dtype=np.uint64)
    fn = lib().okcma_sort_index
    fn.restype, fn.argtypes = None, [_dp, C.c_uint64, C.POINTER(C.c_uint64)]
    fn(_as_dp(f), f.size, out.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out


def eigen(c):
    c = np.ascontiguousarray(c, dtype=np.float64)
    n = c.shape[0]
    w, q = np.empty(n), np.empty((n, n))
    fn = lib().okcma_eigen
    fn.restype, fn.argtypes = C.c_int, [C.c_uint64, _dp, _dp, _dp]
    rc = fn(n, _as_dp(c), _as_dp(w), _as_dp(q))
    assert rc == 0
    return w, q


def sample(z, b, d, mean, sigma):
    z = np.ascontiguousarray(z, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    d = np.ascontiguousarray(d, dtype=np.float64); mean = np.ascontiguousarray(mean, dtype=np.float64)
    rows, n = z.shape
    y, x = np.empty((rows, n)), np.empty((rows, n))
    fn = lib().okcma_sample
    fn.restype, fn.argtypes = None, [C.c_uint64, C.c_uint64, _dp, _dp, _dp, _dp, C.c_double, _dp, _dp]
    fn(n, rows, _as_dp(z), _as_dp(b), _as_dp(d), _as_dp(mean), float(sigma), _as_dp(y), _as_dp(x))
    return y, x


def rank_mu(t, w):
    t = np.ascontiguousarray(t, dtype=np.float64); w = np.ascontiguousarray(w, dtype=np.float64)
    rows, n = t.shape
    p = np.empty((n, n))
    fn = lib().okcma_rank_mu
    fn.restype, fn.argtypes = None, [C.c_uint64, C.c_uint64, _dp, _dp, _dp]
    fn(n, rows, _as_dp(t), _as_dp(w), _as_dp(p))
    return p


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    fn = lib().okcma_philox4x32_10
    fn.restype, fn.argtypes = None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    fn(c, k, o)
    return list(o)


def philox_normal(seed, generation, row_begin, rows, n):
    out = np.empty((rows, n))
    fn = lib().okcma_philox_normal
    fn.restype, fn.argtypes = None, [C.c_uint64] * 5 + [_dp]
    fn(seed, generation, row_begin, rows, n, _as_dp(out))
    return out


def objective(obj, x, coef=None):
    from korali_b200._abi import OBJECTIVES
    x = np.ascontiguousarray(x, dtype=np.float64)
    rows, n = x.shape
    if coef is None:
        coef = 10.0 ** (6.0 * np.arange(n) / max(n - 1, 1))
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    f = np.empty(rows)
    fn = lib().okcma_objective
    fn.restype, fn.argtypes = None, [C.c_int, C.c_uint64, C.c_uint64, _dp, _dp, _dp]
    fn(OBJECTIVES[obj] if isinstance(obj, str) else obj, n, rows, _as_dp(x), _as_dp(coef), _as_dp(f))
    return f


def objective_gradient(obj, x, coef=None):
    from korali_b200._abi import OBJECTIVES
    x = np.ascontiguousarray(x, dtype=np.float64)
    rows, n = x.shape
    if coef is None:
        coef = 10.0 ** (6.0 * np.arange(n) / max(n - 1, 1))
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    g = np.empty((rows, n))
    fn = lib().okcma_objective_gradient
    fn.restype, fn.argtypes = None, [C.c_int, C.c_uint64, C.c_uint64, _dp, _dp, _dp]
    fn(OBJECTIVES[obj] if isinstance(obj, str) else obj, n, rows, _as_dp(x), _as_dp(coef), _as_dp(g))
    return g


def mt19937_gaussian(seed, count, skip=0):
    out = np.empty(count)
    fn = lib().okcma_mt19937_gaussian
    fn.restype, fn.argtypes = None, [C.c_uint64, C.c_uint64, C.c_uint64, _dp]
    fn(seed, skip, count, _as_dp(out))
    return out


def OracleMOCMA(**kw):
    """The MOCMAES oracle (oracle/omocma.c) behind the vocabulary of korali_b200._mocma.MocmaHandle."""
    from korali_b200._mocma import MocmaHandle
    kw.pop("device", None)
    return MocmaHandle(lib(), "omocma_", **kw)

