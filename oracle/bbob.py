"""BBOB objective functions restated on the CPU (oracle, test infrastructure).

The reference evaluates `problem(x)` through ioh 0.3.18 (C++, not installed; call sites
Algorithms/BayesianOptimization/PCA_BO.py:263 and AbstractBayesianOptimizer.py:163). ioh's BBOB
suite is a port of the COCO "bbob2009" legacy code; this file restates the published COCO
definitions (instance data from the legacy LCG `unif`/`gauss`/`compute_rotation`,
`rseed = fid + 10000 * instance`).

Pinned: f15 (RastriginRotated) and f20 (Schwefel) against every row of the reference's shipped
  pca-experiment/data_f15_RastriginRotated/IOHprofiler_f15_DIM5.dat, .../data_f20_Schwefel/...dat
  and the full-precision `best` entries of pca-experiment/IOHprofiler_f15_RastriginRotated.json.
Parity unpinned: f16, f17, f21, f22, f23, f24 (the reference ships no vectors for them).
"""
import math
import numpy as np


# ----------------------------------------------------------------------------- legacy RNG
def unif(n: int, inseed: int) -> np.ndarray:
    """bbob2009_unif: Park-Miller LCG with a 32-entry shuffle table."""
    if inseed < 0:
        inseed = -inseed
    if inseed < 1:
        inseed = 1
    aktseed = inseed
    rgrand = [0] * 32
    for i in range(39, -1, -1):
        tmp = int(math.floor(aktseed / 127773.0))
        aktseed = 16807 * (aktseed - tmp * 127773) - 2836 * tmp
        if aktseed < 0:
            aktseed += 2147483647
        if i < 32:
            rgrand[i] = aktseed
    aktrand = rgrand[0]
    r = np.empty(n)
    for i in range(n):
        tmp = int(math.floor(aktseed / 127773.0))
        aktseed = 16807 * (aktseed - tmp * 127773) - 2836 * tmp
        if aktseed < 0:
            aktseed += 2147483647
        tmp = int(math.floor(aktrand / 67108865.0))
        aktrand = rgrand[tmp]
        rgrand[tmp] = aktseed
        r[i] = aktrand / 2.147483647e9
        if r[i] == 0.0:
            r[i] = 1e-99
    return r


def gauss(n: int, seed: int) -> np.ndarray:
    """bbob2009_gauss: Box-Muller on 2n uniforms."""
    u = unif(2 * n, seed)
    g = np.sqrt(-2.0 * np.log(u[:n])) * np.cos(2.0 * math.pi * u[n:2 * n])
    g[g == 0.0] = 1e-99
    return g


def compute_rotation(seed: int, dim: int) -> np.ndarray:
    """bbob2009_compute_rotation: Gram-Schmidt on the columns of a Gaussian matrix."""
    g = gauss(dim * dim, seed)
    B = g.reshape(dim, dim).T.copy()  # B[i][j] = g[j*dim + i]
    for i in range(dim):
        for j in range(i):
            prod = float(np.dot(B[:, i], B[:, j]))
            B[:, i] -= prod * B[:, j]
        prod = float(np.dot(B[:, i], B[:, i]))
        B[:, i] /= math.sqrt(prod)
    return B


def compute_xopt(seed: int, dim: int) -> np.ndarray:
    x = unif(dim, seed)
    x = 8.0 * np.floor(1e4 * x) / 1e4 - 4.0
    x[x == 0.0] = -1e-5
    return x


def _round_half_away(x: float) -> float:
    return math.floor(x + 0.5)


def compute_fopt(fid: int, instance: int) -> float:
    rseed = {4: 3, 18: 17}.get(fid, fid)
    rrseed = rseed + 10000 * instance
    gval = gauss(1, rrseed)[0]
    gval2 = gauss(1, rrseed + 1)[0]
    return min(1000.0, max(-1000.0, _round_half_away(100.0 * 100.0 * gval / gval2) / 100.0))


# ----------------------------------------------------------------------------- variable transforms
def t_osz(x: np.ndarray) -> np.ndarray:
    """transform_vars_oscillate (alpha = 0.1), elementwise."""
    out = np.zeros_like(x)
    pos = x > 0
    neg = x < 0
    with np.errstate(all="ignore"):
        tp = np.log(np.where(pos, x, 1.0)) / 0.1
        bp = np.exp(tp + 0.49 * (np.sin(tp) + np.sin(0.79 * tp)))
        tn = np.log(np.where(neg, -x, 1.0)) / 0.1
        bn = np.exp(tn + 0.49 * (np.sin(0.55 * tn) + np.sin(0.31 * tn)))
    out = np.where(pos, np.power(bp, 0.1), out)
    out = np.where(neg, -np.power(bn, 0.1), out)
    return out


def t_asy(x: np.ndarray, beta: float) -> np.ndarray:
    """transform_vars_asymmetric."""
    d = x.shape[-1]
    expo = 1.0 + beta * (np.arange(d) / (d - 1.0)) * np.sqrt(np.where(x > 0, x, 0.0))
    return np.where(x > 0, np.power(np.where(x > 0, x, 1.0), expo), x)


def f_pen(x: np.ndarray) -> np.ndarray:
    t = np.abs(x) - 5.0
    t = np.where(t > 0, t, 0.0)
    return np.sum(t * t, axis=-1)


class BBOBProblem:
    """Callable BBOB problem: `p(x)` for one point (float) or `p.batch(X)` for rows of X.

    Mirrors the part of the ioh RealSingleObjective protocol that the reference touches
    (Algorithms/AbstractAlgorithm.py:73-83): `.dim`, `.bounds` ([-5, 5]^D), minimisation.
    """

    NAMES = {15: "RastriginRotated", 16: "Weierstrass", 17: "Schaffers10", 20: "Schwefel", 21: "Gallagher101",
             22: "Gallagher21", 23: "Katsuura", 24: "LunacekBiRastrigin"}

    def __init__(self, fid: int, dim: int, instance: int):
        if fid not in self.NAMES:
            raise ValueError(f"BBOB f{fid} is not restated in the oracle")
        self.fid, self.dim, self.instance = fid, dim, instance
        self.rseed = fid + 10000 * instance
        self.fopt = compute_fopt(fid, instance)
        self.bounds = np.tile(np.array([[-5.0, 5.0]]), (dim, 1))
        self.maximization = False
        self.evaluations = 0
        D = dim
        idx = np.arange(D) / (D - 1.0) if D > 1 else np.zeros(1)
        if fid in (15, 16, 17):
            self.xopt = compute_xopt(self.rseed, D)
            self.rot1 = compute_rotation(self.rseed + 1000000, D)
            self.rot2 = compute_rotation(self.rseed, D)
        if fid == 15:
            lam = np.power(math.sqrt(10.0), idx)
            self.M = (self.rot1 * lam[None, :]) @ self.rot2
        elif fid == 16:
            lam = np.power(1.0 / math.sqrt(100.0), idx)
            self.M = (self.rot1 * lam[None, :]) @ self.rot2
            k = np.arange(12)
            self.ak = np.power(0.5, k)
            self.bk = np.power(3.0, k)
            self.f0 = float(np.sum(self.ak * np.cos(2.0 * math.pi * self.bk * 0.5)))
        elif fid == 17:
            lam = np.power(math.sqrt(10.0), idx)
            self.M = lam[:, None] * self.rot2
        elif fid == 20:
            u = unif(D, self.rseed)
            self.sign = np.where(u - 0.5 < 0, -1.0, 1.0)
            self.xopt = self.sign * 0.5 * 4.2096874637
            self.lam = np.power(math.sqrt(10.0), idx)
        elif fid in (21, 22):
            # COCO legacy f21 / f22 (Gallagher's Gaussian peaks): 101 peaks, first condition sqrt(1000), optimum
            # in 0.8 [-5, 5]; 21 peaks, first condition 1000, optimum in 0.8 [-4.9, 4.9]. One global rotation.
            npk = 101 if fid == 21 else 21
            bb, cc = (10.0, 5.0) if fid == 21 else (9.8, 4.9)
            maxcond = 1000.0
            self.rot = compute_rotation(self.rseed, D)
            rperm = np.argsort(unif(npk - 1, self.rseed), kind="stable")
            cond = np.empty(npk)
            cond[0] = math.sqrt(maxcond) if fid == 21 else maxcond
            self.peakvalues = np.empty(npk)
            self.peakvalues[0] = 10.0
            for i in range(1, npk):
                cond[i] = math.pow(maxcond, rperm[i - 1] / (npk - 2.0))
                self.peakvalues[i] = (i - 1) / (npk - 2.0) * (9.1 - 1.1) + 1.1
            self.scales = np.empty((npk, D))
            for i in range(npk):
                rp = np.argsort(unif(D, self.rseed + 1000 * i), kind="stable")
                self.scales[i] = np.power(cond[i], rp / (D - 1.0) - 0.5) if D > 1 else 1.0
            pk = unif(D * npk, self.rseed)
            self.xopt = 0.8 * (bb * pk[:D] - cc)
            self.xlocal = (bb * pk.reshape(npk, D) - cc) @ self.rot.T       # [peak][coordinate], rotated
            self.xlocal[0] *= 0.8
        elif fid == 23:
            # Katsuura: z = R1 diag(sqrt(100)^(i/(D-1))) R2 (x - xopt)
            self.xopt = compute_xopt(self.rseed, D)
            rot1, rot2 = compute_rotation(self.rseed + 1000000, D), compute_rotation(self.rseed, D)
            self.M = (rot1 * np.power(math.sqrt(100.0), idx)[None, :]) @ rot2
        elif fid == 24:
            # Lunacek bi-Rastrigin: xopt = +-mu1/2 by the sign of a Gaussian draw, same conditioning map as f23
            g = gauss(D, self.rseed)
            self.sign = np.where(g < 0, -1.0, 1.0)
            self.xopt = 0.5 * 2.5 * self.sign
            rot1, rot2 = compute_rotation(self.rseed + 1000000, D), compute_rotation(self.rseed, D)
            self.M = (rot1 * np.power(math.sqrt(100.0), idx)[None, :]) @ rot2

    # ------------------------------------------------------------------ evaluation
    def batch(self, X: np.ndarray) -> np.ndarray:
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        D = self.dim
        if self.fid == 15:
            z = (X - self.xopt) @ self.rot1.T
            z = t_asy(t_osz(z), 0.2)
            z = z @ self.M.T
            y = 10.0 * (D - np.sum(np.cos(2.0 * math.pi * z), axis=-1)) + np.sum(z * z, axis=-1)
        elif self.fid == 16:
            z = (X - self.xopt) @ self.rot1.T
            z = t_osz(z)
            z = z @ self.M.T
            s = np.sum(np.cos(2.0 * math.pi * (z[..., None] + 0.5) * self.bk) * self.ak, axis=(-1, -2))
            y = 10.0 * np.power(s / D - self.f0, 3) + (10.0 / D) * f_pen(X)
        elif self.fid == 17:
            z = (X - self.xopt) @ self.rot1.T
            z = t_asy(z, 0.5)
            z = z @ self.M.T
            s = z[:, :-1] ** 2 + z[:, 1:] ** 2
            r = np.sum(np.power(s, 0.25) * (1.0 + np.sin(50.0 * np.power(s, 0.1)) ** 2), axis=-1)
            y = np.power(r / (D - 1.0), 2) + 10.0 * f_pen(X)
        elif self.fid in (21, 22):
            tmx = X @ self.rot.T
            d = tmx[:, None, :] - self.xlocal[None, :, :]
            q = np.sum(self.scales[None] * d * d, axis=-1)                    # [points, peaks]
            f = 10.0 - np.max(self.peakvalues[None] * np.exp((-0.5 / D) * q), axis=-1)
            y = t_osz(f) ** 2 + f_pen(X)
        elif self.fid == 23:
            z = (X - self.xopt) @ self.M.T
            p2 = np.power(2.0, np.arange(1, 33))
            arr = z[..., None] * p2
            rnd = np.sign(arr) * np.floor(np.abs(arr) + 0.5)                  # C round(): halves away from zero
            t = np.sum(np.abs(arr - rnd) / p2, axis=-1)
            prod = np.prod(1.0 + t * np.arange(1, D + 1), axis=-1)
            y = 10.0 / D / D * (-1.0 + np.power(prod, 10.0 / math.pow(D, 1.2))) + f_pen(X)
        elif self.fid == 24:
            mu1, dd = 2.5, 1.0
            sfac = 1.0 - 0.5 / (math.sqrt(D + 20.0) - 4.1)
            mu2 = -math.sqrt((mu1 * mu1 - dd) / sfac)
            tmx = 2.0 * X * self.sign
            s1 = np.sum((tmx - mu1) ** 2, axis=-1)
            s2 = np.sum((tmx - mu2) ** 2, axis=-1)
            z = (tmx - mu1) @ self.M.T
            y = np.minimum(s1, dd * D + sfac * s2) + 10.0 * (D - np.sum(np.cos(2.0 * math.pi * z), axis=-1)) \
                + 1e4 * f_pen(X)
        else:  # 20
            xh = 2.0 * self.sign * X
            zh = xh.copy()
            zh[:, 1:] = xh[:, 1:] + 0.25 * (xh[:, :-1] - 2.0 * np.abs(self.xopt[:-1]))
            z = 100.0 * (self.lam * (zh - 2.0 * np.abs(self.xopt)) + 2.0 * np.abs(self.xopt))
            t = np.abs(z) - 500.0
            t = np.where(t > 0, t, 0.0)
            pen = np.sum(t * t, axis=-1)
            s = np.sum(z * np.sin(np.sqrt(np.abs(z))), axis=-1)
            y = 0.01 * (pen + 418.9828872724339 - s / D)
        return y + self.fopt

    def __call__(self, x) -> float:
        self.evaluations += 1
        return float(self.batch(np.asarray(x, dtype=np.float64).reshape(1, -1))[0])
