"""CPU test: the team-parallel L-BFGS-B core (host build, one thread) against scipy's L-BFGS-B on box-constrained
test problems -- same evaluation points, iteration and evaluation counts (tests/native/lbfgsb_host.cpp)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
from scipy.optimize import minimize, rosen, rosen_der

HERE = os.path.dirname(os.path.abspath(__file__))
OBJ = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                       ctypes.c_void_p)
DP = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("lb") / "liblbfgsb_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-o", so,
                           os.path.join(HERE, "native", "lbfgsb_host.cpp")])
    l = ctypes.CDLL(so)
    l.lbfgsb_host_minimize.argtypes = [ctypes.c_int, DP, DP, DP, OBJ, ctypes.c_void_p, ctypes.c_int, ctypes.c_double,
                                       ctypes.c_double, DP, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    return l


def run_core(lib, fun, x0, lo, hi, maxiter=200):
    n = len(x0)
    hist = []

    def cb(n_, xp, gp, user):
        x = np.ctypeslib.as_array(xp, (n,)).copy()
        f, g = fun(x)
        hist.append(x)
        np.ctypeslib.as_array(gp, (n,))[:] = g
        return float(f)

    c = OBJ(cb)
    x = np.array(x0, dtype=np.float64)
    lo, hi = np.array(lo, float), np.array(hi, float)
    f, nit, nfev = ctypes.c_double(), ctypes.c_int(), ctypes.c_int()
    P = lambda a: a.ctypes.data_as(DP)  # noqa: E731
    st = lib.lbfgsb_host_minimize(n, P(x), P(lo), P(hi), c, None, maxiter, 1e7, 1e-5, ctypes.byref(f),
                                  ctypes.byref(nit), ctypes.byref(nfev))
    return x, f.value, nit.value, nfev.value, st, hist


def problems():
    rng = np.random.RandomState(0)
    out = []
    for trial in range(12):
        n = [5, 12, 40, 100, 200][trial % 5]
        if trial % 3 == 0:
            fun = lambda x: (rosen(x), rosen_der(x))  # noqa: E731
            lo, hi, x0 = np.full(n, -1.5), np.full(n, 1.2), rng.uniform(-2, 2, n)
        elif trial % 3 == 1:
            Q = rng.randn(n, n)
            A, b = Q @ Q.T + 0.1 * np.eye(n), rng.randn(n) * 3
            fun = lambda x, A=A, b=b: (0.5 * x @ A @ x - b @ x, A @ x - b)  # noqa: E731
            lo, hi, x0 = np.full(n, -0.5), np.full(n, 0.7), rng.uniform(-1, 1, n)
        else:
            c = rng.randn(n)
            fun = lambda x, c=c: (np.sum(np.sin(3 * x + c)) + 0.1 * np.sum(x ** 4), 3 * np.cos(3 * x + c) + 0.4 * x ** 3)  # noqa: E731
            lo, hi, x0 = np.full(n, -2.0), np.full(n, 2.0), rng.uniform(-2, 2, n)
        out.append((n, fun, np.clip(x0, lo, hi), lo, hi))
    return out


def test_core_follows_scipy(lib):
    exact = 0
    for n, fun, x0, lo, hi in problems():
        hs = []

        def f_sp(x):
            hs.append(x.copy())
            return fun(x)

        res = minimize(f_sp, x0, jac=True, method="L-BFGS-B", bounds=list(zip(lo, hi)), options={"maxiter": 200})
        x, f, nit, nfev, st, hist = run_core(lib, fun, x0, lo, hi)
        assert np.all(x >= lo) and np.all(x <= hi)
        if res.status == 0:
            # converged runs: identical evaluation sequence (to rounding), iteration and evaluation counts
            assert st in (0, 1) and nit == res.nit and nfev == res.nfev, (n, nit, res.nit, nfev, res.nfev)
            dev = max(np.max(np.abs(a - b)) for a, b in zip(hs, hist))
            assert dev < 1e-7, (n, dev)
            assert np.max(np.abs(x - res.x)) < 1e-9 and abs(f - res.fun) <= 1e-9 * max(1.0, abs(res.fun))
            exact += 1
        else:
            # runs cut by maxiter = 200 are chaotic in rounding: same stop reason, first 40 evaluations identical
            assert st == 2 and nit == res.nit == 200
            assert max(np.max(np.abs(a - b)) for a, b in zip(hs[:40], hist[:40])) < 1e-6
    assert exact >= 9


def test_core_start_at_optimum_and_fixed_variables(lib):
    # start at a KKT point: converges immediately on the projected gradient
    fun = lambda x: (float(np.sum((x - 3.0) ** 2)), 2 * (x - 3.0))  # noqa: E731
    x, f, nit, nfev, st, _ = run_core(lib, fun, np.full(4, 1.0), np.full(4, -1.0), np.full(4, 1.0))
    assert st == 0 and nit == 0 and nfev == 1 and np.all(x == 1.0)
    # all variables run into their bounds
    x, f, nit, nfev, st, _ = run_core(lib, fun, np.zeros(4), np.full(4, -1.0), np.full(4, 1.0))
    assert np.all(x == 1.0) and st in (0, 1)
