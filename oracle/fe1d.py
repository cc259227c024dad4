"""1D finite-element data for FE_Q(k) with QGauss(k+1) on the unit interval.

Oracle restatement (test infrastructure, see oracle/__init__.py) of what
deal.II's ``ShapeInfo`` provides to the reference's evaluators:
 * FE_Q(k) support points are the Gauss-Lobatto points on [0,1]
   (reference builds ``FE_Q<dim> fe(degree)``: benchmark_01.h:244,
   benchmark_03.h:434);
 * quadrature is ``QGauss(k+1)`` (benchmark_01.h:245, benchmark_03.h:435);
 * ``subface_interpolation_matrices[s][i][j] = l_j((x_i + s)/2)`` is the
   interpolation from a coarse edge to its lower (s=0) / upper (s=1) half,
   which ``FEEvaluationHangingNodesFactory::apply`` consumes
   (benchmark_00_likwid.cc:56-59).

All arrays are computed in extended precision (numpy longdouble) and rounded
to float64 once.
"""
from __future__ import annotations

import functools

import numpy as np

LD = np.longdouble


def _legendre(n: int, x):
    """P_n(x) and P_n'(x) on [-1,1] by the three-term recurrence (longdouble)."""
    x = np.asarray(x, dtype=LD)
    p0 = np.ones_like(x)
    if n == 0:
        return p0, np.zeros_like(x)
    p1 = x.copy()
    for m in range(2, n + 1):
        p0, p1 = p1, ((2 * m - 1) * x * p1 - (m - 1) * p0) / m
    dp = n * (x * p1 - p0) / (x * x - 1)
    return p1, dp


@functools.lru_cache(maxsize=None)
def gauss(n: int):
    """n-point Gauss-Legendre points and weights on [0,1] (longdouble)."""
    i = np.arange(1, n + 1, dtype=LD)
    x = -np.cos(np.pi * (i - LD(0.25)) / (n + LD(0.5)))
    for _ in range(100):
        p, dp = _legendre(n, x)
        dx = p / dp
        x = x - dx
        if np.max(np.abs(dx)) < 1e-19:
            break
    _, dp = _legendre(n, x)
    w = 2 / ((1 - x * x) * dp * dp)
    x = (x + 1) / 2
    w = w / 2
    # enforce exact symmetry
    x = (x + (1 - x[::-1])) / 2
    w = (w + w[::-1]) / 2
    return x, w


@functools.lru_cache(maxsize=None)
def gauss_lobatto(n: int):
    """n-point Gauss-Lobatto points on [0,1] (longdouble), n >= 2."""
    if n == 2:
        return np.array([0, 1], dtype=LD)
    m = n - 1  # interior points are the roots of P_m'
    # Chebyshev-Gauss-Lobatto initial guess for the interior nodes
    j = np.arange(1, m, dtype=LD)
    x = -np.cos(np.pi * j / m)
    for _ in range(100):
        p, dp = _legendre(m, x)
        # P_m'' from the Legendre ODE: (1-x^2) P'' = 2x P' - m(m+1) P
        d2p = (2 * x * dp - m * (m + 1) * p) / (1 - x * x)
        dx = dp / d2p
        x = x - dx
        if np.max(np.abs(dx)) < 1e-19:
            break
    x = np.concatenate([[LD(-1)], x, [LD(1)]])
    x = (x + 1) / 2
    x = (x + (1 - x[::-1])) / 2
    return x


def lagrange(nodes, pts):
    """Values V[q][j] = l_j(pts[q]) and derivatives D[q][j] = l_j'(pts[q])
    of the Lagrange basis on ``nodes`` (longdouble, direct product formula)."""
    nodes = np.asarray(nodes, dtype=LD)
    pts = np.asarray(pts, dtype=LD)
    n = len(nodes)
    V = np.ones((len(pts), n), dtype=LD)
    D = np.zeros((len(pts), n), dtype=LD)
    for j in range(n):
        others = [m for m in range(n) if m != j]
        denom = LD(1)
        for m in others:
            denom *= nodes[j] - nodes[m]
        for q, x in enumerate(pts):
            v = LD(1)
            for m in others:
                v *= x - nodes[m]
            V[q, j] = v / denom
            d = LD(0)
            for m in others:
                t = LD(1)
                for r in others:
                    if r != m:
                        t *= x - nodes[r]
                d += t
            D[q, j] = d / denom
    return V, D


class ShapeData:
    """Everything the operator needs in 1D for degree k (float64 arrays).

    nodes      GLL support points x_i on [0,1]            (k+1)
    qpts, qw   Gauss points / weights on [0,1]            (k+1)
    S[q][i]    l_i(qpts[q])        nodal basis -> values at Gauss points
    G[q][i]    l_i'(qpts[q])       nodal basis -> derivative at Gauss points
    Dc[q][p]   derivative of the Gauss-collocation Lagrange basis at Gauss pts
               (S then Dc equals G: the collocation-gradient factorisation the
               CPU evaluator uses)
    W[s][i][j] subface interpolation l_j((x_i+s)/2), s = 0 lower, 1 upper
    """

    def __init__(self, degree: int):
        assert degree >= 1
        self.degree = degree
        n = degree + 1
        self.n = n
        x = gauss_lobatto(n)
        q, w = gauss(n)
        S, G = lagrange(x, q)
        _, Dc = lagrange(q, q)
        W0, _ = lagrange(x, x / 2)
        W1, _ = lagrange(x, (x + 1) / 2)
        self._ld = dict(nodes=x, qpts=q, qw=w, S=S, G=G, Dc=Dc, W=np.stack([W0, W1]))
        self.nodes = x.astype(np.float64)
        self.qpts = q.astype(np.float64)
        self.qw = w.astype(np.float64)
        self.S = S.astype(np.float64)
        self.G = G.astype(np.float64)
        self.Dc = Dc.astype(np.float64)
        self.W = np.stack([W0, W1]).astype(np.float64)

    def longdouble(self, name):
        return self._ld[name]

    def eval_basis(self, xi):
        """Nodal basis values at arbitrary reference points xi (float64)."""
        V, _ = lagrange(self._ld["nodes"], np.atleast_1d(xi))
        return V.astype(np.float64)


@functools.lru_cache(maxsize=None)
def shape_data(degree: int) -> ShapeData:
    return ShapeData(degree)
