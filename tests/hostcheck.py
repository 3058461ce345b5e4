"""ctypes access to libpyro_b200_hostcheck.so: the __host__ __device__ functors of
pyro_b200/csrc compiled for the CPU.  TEST INFRASTRUCTURE ONLY (nothing in pyro_b200/ loads it):
it lets the CPU test tier pin the exact arithmetic the kernels execute, and the NUTS tree code,
against the oracle before any GPU time is spent."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(os.path.join(ROOT, "pyro_b200", "libpyro_b200_hostcheck.so"))
        _lib.b2h_digamma.restype = ctypes.c_double
        _lib.b2h_digamma.argtypes = [ctypes.c_double]
        _lib.b2h_trigamma.restype = ctypes.c_double
        _lib.b2h_trigamma.argtypes = [ctypes.c_double]
        _lib.b2h_potential_hier_normal.restype = ctypes.c_double
        _lib.b2h_potential_logistic.restype = ctypes.c_double
    return _lib


def _ptr(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct)) if a is not None else None


def eval_family(fam, x, params, dtype=np.float64, grad=True):
    n = len(x) if x is not None else len(params[0])
    ct = ctypes.c_double if dtype == np.float64 else ctypes.c_float
    P = ctypes.POINTER(ct)
    xs = None if x is None else np.ascontiguousarray(x, dtype=dtype)
    pa = [np.ascontiguousarray(p, dtype=dtype) for p in params] + [None] * (4 - len(params))
    lp = np.zeros(n, dtype)
    dx = np.zeros(n, dtype)
    dp = [np.zeros(n, dtype) for _ in range(4)]
    pp = (P * 4)(*[_ptr(a, ct) for a in pa])
    dpp = (P * 4)(*[_ptr(a, ct) for a in dp])
    f = lib().b2h_eval_f64 if dtype == np.float64 else lib().b2h_eval_f32
    rc = f(fam, int(grad), ctypes.c_int64(n), _ptr(xs, ct), pp, _ptr(lp, ct), _ptr(dx, ct), dpp)
    assert rc == 0
    return lp, dx, dp[: len(params)]


def potential_hier_normal(y, sigma, s_mu, s_tau, z):
    y = np.ascontiguousarray(y, np.float64); sigma = np.ascontiguousarray(sigma, np.float64)
    z = np.ascontiguousarray(z, np.float64)
    g = np.zeros_like(z)
    d = ctypes.c_double
    U = lib().b2h_potential_hier_normal(_ptr(y, d), _ptr(sigma, d), ctypes.c_int64(len(y)),
                                        d(s_mu), d(s_tau), _ptr(z, d), _ptr(g, d))
    return U, g


def potential_logistic(X, y, s, z):
    X = np.ascontiguousarray(X, np.float64); y = np.ascontiguousarray(y, np.float64)
    z = np.ascontiguousarray(z, np.float64)
    g = np.zeros_like(z)
    d = ctypes.c_double
    U = lib().b2h_potential_logistic(_ptr(X, d), _ptr(y, d), ctypes.c_int64(X.shape[0]),
                                     ctypes.c_int(X.shape[1]), d(s), _ptr(z, d), _ptr(g, d))
    return U, g


def nuts_hier_normal(y, sigma, s_mu, s_tau, z, U, g, eps, minv, T, max_depth=10, seed=0):
    d = ctypes.c_double
    y = np.ascontiguousarray(y, np.float64); sigma = np.ascontiguousarray(sigma, np.float64)
    C, D = z.shape
    samples = np.zeros((T, C, D)); acc = np.zeros((T, C))
    depth = np.zeros((T, C), np.int32); div = np.zeros((T, C), np.int32); steps = np.zeros((T, C), np.int32)
    i32 = ctypes.c_int32
    rc = lib().b2h_nuts_hier_normal(_ptr(y, d), _ptr(sigma, d), ctypes.c_int64(len(y)), d(s_mu), d(s_tau),
                                    ctypes.c_int64(C), _ptr(z, d), _ptr(U, d), _ptr(g, d), _ptr(eps, d),
                                    _ptr(minv, d), ctypes.c_int(T), ctypes.c_int(max_depth), d(1000.0),
                                    ctypes.c_uint64(seed), _ptr(samples, d), _ptr(acc, d),
                                    _ptr(depth, i32), _ptr(div, i32), _ptr(steps, i32))
    assert rc == 0
    return samples, acc, depth, div, steps
