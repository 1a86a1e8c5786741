"""Minimal CMA-ES with the slice of pycma's interface that the reference's host loop uses
(st_ito/style_transfer.py:614-673): ``CMAEvolutionStrategy(x0, sigma0, {"bounds": [0, 1],
"popsize": P})``, ``ask()``, ``tell(X, fvals)``, ``result[0] / result[1]``, ``disp()``, ``stop()``.

pycma is not a dependency of this repo (and is absent offline), so this is a standard
(mu/mu_w, lambda)-CMA-ES (Hansen, "The CMA Evolution Strategy: A Tutorial") with rank-one and
rank-mu updates, cumulative step-size adaptation and pycma-style smooth box-boundary handling (a
piecewise linear/quadratic, periodic map of the search point into the box, so every candidate handed
to the objective is feasible).  It takes an explicit ``seed`` (the reference passes none and is not
reproducible); with the same seed every rank of a multi-GPU run draws the same population, so no
broadcast of W is needed.  Sampling differs from pycma (no active/negative weights); parity of this
repo is defined on evaluate(W) for a given W, which is sampler-independent.
"""
from __future__ import annotations

import math
from collections import namedtuple

import numpy as np

CMAResult = namedtuple("CMAResult", ["xbest", "fbest", "evals_best", "evaluations", "iterations", "xfavorite",
                                     "stds", "stop"])


class _BoxTransform:
    """Smooth map R -> [lb, ub]: identity inside [lb+al, ub-au], quadratic near the bounds,
    mirrored/periodic outside (same construction as pycma's BoxConstraintsLinQuadTransformation)."""

    def __init__(self, lb, ub, n):
        self.lb = np.full(n, float(lb))
        self.ub = np.full(n, float(ub))
        span = self.ub - self.lb
        self.al = np.minimum(span / 2.0, (1.0 + np.abs(self.lb)) / 20.0)
        self.au = np.minimum(span / 2.0, (1.0 + np.abs(self.ub)) / 20.0)

    def __call__(self, x):
        x = np.array(x, dtype=np.float64, copy=True)
        lb, ub, al, au = self.lb, self.ub, self.al, self.au
        span = ub - lb
        far = (x < lb - 2 * al - span / 2.0) | (x > ub + 2 * au + span / 2.0)
        if far.any():
            r = 2 * (span + al + au)
            s = lb - 2 * al - span / 2.0
            x = np.where(far, x - r * np.floor((x - s) / r), x)
        x = np.where(x > ub + au, x - 2 * (x - ub - au), x)
        x = np.where(x < lb - al, x + 2 * (lb - al - x), x)
        lo = x < lb + al
        hi = (~lo) & (x >= ub - au)
        y = x.copy()
        y[lo] = (lb + (x - (lb - al)) ** 2 / 4.0 / al)[lo]
        y[hi] = (ub - (x - (ub + au)) ** 2 / 4.0 / au)[hi]
        return np.clip(y, lb, ub)

    def inverse(self, y):
        """Pre-image in [lb - al, ub + au] of a point of the box (pycma: gp.geno(x0) starts the search at the
        genotype whose phenotype is x0, so an x0 within al / au of a bound is not shifted by the first ask())."""
        y = np.clip(np.array(y, dtype=np.float64, copy=True), self.lb, self.ub)
        lb, ub, al, au = self.lb, self.ub, self.al, self.au
        x = y.copy()
        lo = y < lb + al
        hi = (~lo) & (y > ub - au)
        x[lo] = ((lb - al) + 2.0 * np.sqrt(al * (y - lb)))[lo]
        x[hi] = ((ub + au) - 2.0 * np.sqrt(au * (ub - y)))[hi]
        return x


class PyCMAEvolutionStrategy:
    """Pure-numpy implementation (kept as the readable statement of the algorithm and for environments without the
    built library); ``CMAEvolutionStrategy`` below runs the same algorithm natively inside libstito."""

    def __init__(self, x0, sigma0, inopts=None):
        opts = dict(inopts or {})
        self.N = N = int(np.asarray(x0).size)
        self.xmean = np.array(x0, dtype=np.float64).reshape(N)
        self.sigma = float(sigma0)
        self.popsize = int(opts.get("popsize") or (4 + int(3 * math.log(N))))
        self.rng = np.random.RandomState(opts.get("seed", None))
        self.verbose = opts.get("verbose", 1)
        bounds = opts.get("bounds")
        self._box = _BoxTransform(bounds[0], bounds[1], N) if bounds is not None else None
        if self._box is not None:
            self.xmean = self._box.inverse(self.xmean)
        lam = self.popsize
        self.mu = mu = lam // 2
        w = math.log(mu + 0.5) - np.log(np.arange(1, mu + 1))
        self.weights = w / w.sum()
        self.mueff = 1.0 / np.sum(self.weights ** 2)
        me = self.mueff
        self.cc = (4 + me / N) / (N + 4 + 2 * me / N)
        self.cs = (me + 2) / (N + me + 5)
        self.c1 = 2 / ((N + 1.3) ** 2 + me)
        self.cmu = min(1 - self.c1, 2 * (me - 2 + 1 / me) / ((N + 2) ** 2 + me))
        self.damps = 1 + 2 * max(0.0, math.sqrt((me - 1) / (N + 1)) - 1) + self.cs
        self.chiN = math.sqrt(N) * (1 - 1 / (4 * N) + 1 / (21 * N * N))
        self.pc = np.zeros(N)
        self.ps = np.zeros(N)
        self.B = np.eye(N)
        self.D = np.ones(N)
        self.C = np.eye(N)
        self.invsqrtC = np.eye(N)
        self._eigen_at = 0
        self.countevals = 0
        self.countiter = 0
        self._geno = None
        self._best_x = None
        self._best_f = float("inf")
        self._best_evals = 0
        self._last_f = None

    # -- pycma surface -------------------------------------------------------------------------
    def ask(self):
        lam, N = self.popsize, self.N
        z = self.rng.randn(lam, N)
        y = (z * self.D) @ self.B.T
        self._geno = self.xmean + self.sigma * y
        # the box map is element-wise: one vectorised call for the whole population (a Python loop over the
        # candidates cost 5.6 ms per generation at popsize 64 -- 40 % of a B200 generation)
        pheno = self._geno if self._box is None else self._box(self._geno)
        return list(np.array(pheno))

    def tell(self, solutions, function_values):
        f = np.asarray(function_values, dtype=np.float64).reshape(-1)
        if self._geno is None or len(solutions) != self.popsize or f.size != self.popsize:
            raise ValueError("tell() needs the popsize solutions of the preceding ask()")
        f = np.where(np.isfinite(f), f, np.inf)
        N, mu = self.N, self.mu
        self.countevals += self.popsize
        self.countiter += 1
        order = np.argsort(f, kind="stable")
        if f[order[0]] < self._best_f:
            self._best_f = float(f[order[0]])
            self._best_x = np.array(solutions[order[0]], dtype=np.float64, copy=True)
            self._best_evals = self.countevals - self.popsize + int(order[0]) + 1
        self._last_f = f[order]
        xold = self.xmean
        sel = self._geno[order[:mu]]
        self.xmean = self.weights @ sel
        ymean = (self.xmean - xold) / self.sigma
        self.ps = (1 - self.cs) * self.ps + math.sqrt(self.cs * (2 - self.cs) * self.mueff) * (self.invsqrtC @ ymean)
        hsig = (np.linalg.norm(self.ps) / math.sqrt(1 - (1 - self.cs) ** (2 * self.countiter)) / self.chiN
                < 1.4 + 2 / (N + 1))
        self.pc = (1 - self.cc) * self.pc + hsig * math.sqrt(self.cc * (2 - self.cc) * self.mueff) * ymean
        artmp = (sel - xold) / self.sigma
        self.C = ((1 - self.c1 - self.cmu) * self.C
                  + self.c1 * (np.outer(self.pc, self.pc) + (1 - hsig) * self.cc * (2 - self.cc) * self.C)
                  + self.cmu * (artmp.T * self.weights) @ artmp)
        self.sigma *= math.exp((self.cs / self.damps) * (np.linalg.norm(self.ps) / self.chiN - 1))
        if self.countevals - self._eigen_at > self.popsize / (self.c1 + self.cmu) / N / 10:
            self._eigen_at = self.countevals
            self.C = np.triu(self.C) + np.triu(self.C, 1).T
            d2, self.B = np.linalg.eigh(self.C)
            self.D = np.sqrt(np.maximum(d2, 1e-30))
            self.invsqrtC = (self.B / self.D) @ self.B.T
        self._geno = None

    @property
    def result(self):
        fav = self.xmean if self._box is None else self._box(self.xmean)
        return CMAResult(self._best_x, self._best_f, self._best_evals, self.countevals, self.countiter, fav,
                         self.sigma * np.sqrt(np.maximum(np.diag(self.C), 0)), self.stop())

    def stop(self):
        out = {}
        if self.sigma * float(np.max(self.D)) < 1e-11:
            out["tolx"] = 1e-11
        if self._last_f is not None and self.countiter > 10 and float(np.ptp(self._last_f)) < 1e-11:
            out["tolfun"] = 1e-11
        return out

    def disp(self, modulo=None):
        if not self.verbose:
            return
        if self.countiter == 1:
            print("Iterat #Fevals   function value  axis ratio  sigma  min&max std")
        stds = self.sigma * np.sqrt(np.maximum(np.diag(self.C), 0))
        fbest = float(self._last_f[0]) if self._last_f is not None else float("nan")
        print(f"{self.countiter:5d} {self.countevals:7d} {fbest: .15e} {float(self.D.max() / self.D.min()):.1e} "
              f"{self.sigma:.2e}  {stds.min():.0e} {stds.max():.0e}")


class NativeCMAEvolutionStrategy:
    """The same CMA-ES behind the C ABI (st_ito_b200/csrc/cma_host.cpp: stito_cma_*): ask + tell cost ~0.1 ms instead of
    ~0.7 ms of numpy per generation, which matters next to a 2-3 ms sharded B200 generation.  Interface = the slice of
    pycma the reference's host loop uses (style_transfer.py:614-673).  Gaussian draws come from the library's documented
    counter-based generator, so trajectories differ from the numpy class for the same seed (parity of this project is
    defined on evaluate(W) for a given W)."""

    def __init__(self, x0, sigma0, inopts=None):
        from ctypes import byref, c_void_p

        from . import _lib

        opts = dict(inopts or {})
        self._L = _lib.lib()
        x0 = np.ascontiguousarray(np.asarray(x0, dtype=np.float64).reshape(-1))
        self.N = int(x0.size)
        self.popsize = int(opts.get("popsize") or (4 + int(3 * math.log(self.N))))
        self.verbose = opts.get("verbose", 1)
        bounds = opts.get("bounds")
        lo, hi = (float(bounds[0]), float(bounds[1])) if bounds is not None else (0.0, 0.0)
        seed = opts.get("seed", None)
        if seed is None:  # the reference passes no seed: every run differs
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        self._h = c_void_p()
        rc = self._L.stito_cma_create(x0.ctypes.data, self.N, float(sigma0), self.popsize, lo, hi, int(seed) & (2 ** 64 - 1),
                                      byref(self._h))
        if rc != 0:
            raise ValueError(f"stito_cma_create failed ({rc}): x0 [{self.N}], sigma0 {sigma0}, popsize {self.popsize}")
        self._X = np.empty((self.popsize, self.N), dtype=np.float64)
        self._asked = False

    def __del__(self):
        try:
            if self._h:
                self._L.stito_cma_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def ask(self):
        rc = self._L.stito_cma_ask(self._h, self._X.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"stito_cma_ask failed ({rc})")
        self._asked = True
        return list(self._X.copy())

    def tell(self, solutions, function_values):
        f = np.ascontiguousarray(np.asarray(function_values, dtype=np.float64).reshape(-1))
        if not self._asked or len(solutions) != self.popsize or f.size != self.popsize:
            raise ValueError("tell() needs the popsize solutions of the preceding ask()")
        X = np.ascontiguousarray(np.asarray(solutions, dtype=np.float64).reshape(self.popsize, self.N))
        rc = self._L.stito_cma_tell(self._h, X.ctypes.data, f.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"stito_cma_tell failed ({rc})")
        self._asked = False

    def _query(self):
        from ctypes import byref, c_double, c_int, c_int64

        xb, xf, stds = np.empty(self.N), np.empty(self.N), np.empty(self.N)
        fb, sg, ar, fr = c_double(), c_double(), c_double(), c_double()
        hb, eb, ev, it = c_int(), c_int64(), c_int64(), c_int64()
        self._L.stito_cma_result(self._h, xb.ctypes.data, byref(fb), byref(hb), xf.ctypes.data, byref(sg), stds.ctypes.data,
                                 byref(eb), byref(ev), byref(it), byref(ar), byref(fr))
        return xb, fb.value, bool(hb.value), xf, sg.value, stds, eb.value, ev.value, it.value, ar.value, fr.value

    @property
    def result(self):
        xb, fb, hb, xf, sg, stds, eb, ev, it, ar, fr = self._query()
        return CMAResult(xb if hb else None, fb if hb else float("inf"), eb, ev, it, xf, stds, self._stop(sg, stds, it, fr))

    @property
    def sigma(self):
        return self._query()[4]

    @property
    def countiter(self):
        return self._query()[8]

    def _stop(self, sg, stds, it, fr):
        out = {}
        if float(np.max(stds)) < 1e-11:
            out["tolx"] = 1e-11
        if it > 10 and 0 <= fr < 1e-11:
            out["tolfun"] = 1e-11
        return out

    def stop(self):
        xb, fb, hb, xf, sg, stds, eb, ev, it, ar, fr = self._query()
        return self._stop(sg, stds, it, fr)

    def disp(self, modulo=None):
        if not self.verbose or self.verbose < 0:
            return
        xb, fb, hb, xf, sg, stds, eb, ev, it, ar, fr = self._query()
        if it == 1:
            print("Iterat #Fevals   function value  axis ratio  sigma  min&max std")
        print(f"{it:5d} {ev:7d} {fb: .15e} {ar:.1e} {sg:.2e}  {stds.min():.0e} {stds.max():.0e}")


def _native_available() -> bool:
    import os

    from . import _lib

    return os.path.isfile(_lib.LIB_PATH)


class CMAEvolutionStrategy:
    """``cma.CMAEvolutionStrategy`` of the host loop: the native implementation when libstito is built (it always is on
    the product path), the numpy one otherwise or with ``{"implementation": "numpy"}`` in the options."""

    def __new__(cls, x0, sigma0, inopts=None):
        opts = dict(inopts or {})
        impl = opts.pop("implementation", None)
        if impl == "numpy" or (impl is None and not _native_available()):
            return PyCMAEvolutionStrategy(x0, sigma0, opts)
        return NativeCMAEvolutionStrategy(x0, sigma0, opts)
