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


class CMAEvolutionStrategy:
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
