"""Two independently coded CPU Laplace operators -- oracle (test
infrastructure, see oracle/__init__.py; parity unpinned by the reference).

O2  ``vmult_fast``     scalar/numpy restatement of the reference's hot loop
    (benchmark_01.h:579-617 CG(SC) mode; benchmark_03.h:244-270, 297-313):
    gather -> hanging-node interpolation -> evaluate(gradients) ->
    submit_gradient(get_gradient) -> integrate(gradients) ->
    hanging-node interpolation^T -> scatter-add.
O1  ``GeneralOperator`` the general-purpose cross-check the reference sets up
    with ``use_fast_hanging_node_algorithm=false`` +
    ``DoFTools::make_hanging_node_constraints`` (benchmark_01.h:286-293,
    benchmark_02.cc:111-120): dense element matrices on the *raw* cell DoFs
    and an explicit constraint matrix built from geometry only (no masks),
    A = C^T K C.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import fe1d


# --------------------------------------------------------------------------
# O2: fast hanging-node algorithm
# --------------------------------------------------------------------------
def hn_selection(kind: int, d: int, k: int):
    """bool[n,n] over the two transversal directions (t0 < t1, indexed
    [a_t1][a_t0]): which lines along d are interpolated in pass d."""
    n = k + 1
    t0, t1 = [t for t in range(3) if t != d]
    sel = np.zeros((n, n), dtype=bool)
    if kind == 0:
        return sel
    b = [1 - ((kind >> t) & 1) for t in range(3)]  # child bit = 1 - subcell bit
    face = [(kind >> (3 + t)) & 1 for t in range(3)]
    edge = [(kind >> (6 + t)) & 1 for t in range(3)]
    if face[t0]:
        sel[:, b[t0] * k] = True
    if face[t1]:
        sel[b[t1] * k, :] = True
    if edge[d]:
        sel[b[t1] * k, b[t0] * k] = True
    return sel


def hn_apply(u, kinds, degree, transpose):
    """In-place hanging-node interpolation (or its transpose) on cell-local
    values u[c, z, y, x]  (FEEvaluationHangingNodesFactory::apply,
    benchmark_00_likwid.cc:56-59; three directional passes, Appendix B 4-5)."""
    k = degree
    sd = fe1d.shape_data(k)
    for kind in np.unique(kinds):
        kind = int(kind)
        if kind == 0:
            continue
        cells = np.nonzero(kinds == kind)[0]
        v = u[cells]
        for d in range(3):
            sel = hn_selection(kind, d, k)
            if not sel.any():
                continue
            s = 1 - ((kind >> d) & 1)  # subcell bit set => lower child => W_0
            W = sd.W[s].T if transpose else sd.W[s]
            axis = 3 - d  # u is [c,z,y,x]
            vm = np.moveaxis(v, axis, -1)  # [c, t1, t0, line]
            new = vm @ W.T
            vm[:, sel, :] = new[:, sel, :]
        u[cells] = v
    return u


def cell_laplace(u, h, degree):
    """evaluate(gradients) / q-point op / integrate(gradients) on u[c,z,y,x]
    for Cartesian cells of edge length h[c] (benchmark_01.h:603-608)."""
    sd = fe1d.shape_data(degree)
    S, Dc, w = sd.S, sd.Dc, sd.qw
    # basis change to Gauss collocation
    uq = np.einsum("qx,czyx->czyq", S, u)
    uq = np.einsum("qy,czyx->czqx", S, uq)
    uq = np.einsum("qz,czyx->cqyx", S, uq)
    # collocation gradient (reference-cell gradient), q-point factor w_q * h
    w3 = w[:, None, None] * w[None, :, None] * w[None, None, :]
    fac = w3[None] * h[:, None, None, None]
    gx = np.einsum("qx,czyx->czyq", Dc, uq) * fac
    gy = np.einsum("qy,czyx->czqx", Dc, uq) * fac
    gz = np.einsum("qz,czyx->cqyx", Dc, uq) * fac
    r = np.einsum("qx,czyq->czyx", Dc, gx)
    r += np.einsum("qy,czqx->czyx", Dc, gy)
    r += np.einsum("qz,cqyx->czyx", Dc, gz)
    r = np.einsum("qz,cqyx->czyx", S, r)
    r = np.einsum("qy,czqx->czyx", S, r)
    r = np.einsum("qx,czyq->czyx", S, r)
    return r


def cell_laplace_general(u, G, degree):
    """The same evaluate / q-point / integrate sequence with a symmetric 3x3 coefficient per quadrature
    point, G[c, (xx,xy,xz,yy,yz,zz), q] = JxW J^-1 J^-T (lexicographic q): curved cells / high-order mappings
    (TestHighOrderMapping, benchmark_01.h:225-242)."""
    sd = fe1d.shape_data(degree)
    S, Dc = sd.S, sd.Dc
    n = degree + 1
    uq = np.einsum("qx,czyx->czyq", S, u)
    uq = np.einsum("qy,czyx->czqx", S, uq)
    uq = np.einsum("qz,czyx->cqyx", S, uq)
    gx = np.einsum("qx,czyx->czyq", Dc, uq)
    gy = np.einsum("qy,czyx->czqx", Dc, uq)
    gz = np.einsum("qz,czyx->cqyx", Dc, uq)
    Gq = G.reshape(-1, 6, n, n, n)
    hx = Gq[:, 0] * gx + Gq[:, 1] * gy + Gq[:, 2] * gz
    hy = Gq[:, 1] * gx + Gq[:, 3] * gy + Gq[:, 4] * gz
    hz = Gq[:, 2] * gx + Gq[:, 4] * gy + Gq[:, 5] * gz
    r = np.einsum("qx,czyq->czyx", Dc, hx)
    r += np.einsum("qy,czqx->czyx", Dc, hy)
    r += np.einsum("qz,cqyx->czyx", Dc, hz)
    r = np.einsum("qz,cqyx->czyx", S, r)
    r = np.einsum("qy,czqx->czyx", S, r)
    r = np.einsum("qx,czyq->czyx", S, r)
    return r


def vmult_general(lay, src, G, apply_constraints=True, chunk=2048):
    """dst = A src with per-quadrature-point coefficients (fast hanging-node algorithm around it)."""
    n = lay.degree + 1
    dst = np.zeros_like(src)
    for c0 in range(0, lay.n_cells, chunk):
        c1 = min(lay.n_cells, c0 + chunk)
        idx = lay.dof_indices[c0:c1]
        u = src[idx].reshape(-1, n, n, n)
        if apply_constraints:
            hn_apply(u, lay.kinds[c0:c1], lay.degree, transpose=False)
        r = cell_laplace_general(u, G[c0:c1], lay.degree)
        if apply_constraints:
            hn_apply(r, lay.kinds[c0:c1], lay.degree, transpose=True)
        np.add.at(dst, idx.ravel(), r.ravel())
    return dst


def vmult_abs_bound(lay, src, chunk=4096):
    """|A| |src| evaluated through the same pipeline with every 1D matrix replaced by
    its absolute value: the magnitude any floating-point evaluation of A src works
    against (used to scale the single-precision parity check on smooth inputs,
    where A src itself suffers cancellation)."""
    n = lay.degree + 1
    sd = fe1d.shape_data(lay.degree)
    S, Dc, W, w = np.abs(sd.S), np.abs(sd.Dc), np.abs(sd.W), sd.qw
    dst = np.zeros_like(src)
    x = np.abs(src)
    for c0 in range(0, lay.n_cells, chunk):
        c1 = min(lay.n_cells, c0 + chunk)
        idx = lay.dof_indices[c0:c1]
        u = x[idx].reshape(-1, n, n, n)
        kinds = lay.kinds[c0:c1]
        for transpose in (False, True):
            if transpose:
                uq = np.einsum("qx,czyx->czyq", S, u)
                uq = np.einsum("qy,czyx->czqx", S, uq)
                uq = np.einsum("qz,czyx->cqyx", S, uq)
                w3 = w[:, None, None] * w[None, :, None] * w[None, None, :]
                fac = w3[None] * lay.h[c0:c1, None, None, None]
                r = np.einsum("qx,czyq->czyx", Dc, np.einsum("qx,czyx->czyq", Dc, uq) * fac)
                r += np.einsum("qy,czqx->czyx", Dc, np.einsum("qy,czyx->czqx", Dc, uq) * fac)
                r += np.einsum("qz,cqyx->czyx", Dc, np.einsum("qz,czyx->cqyx", Dc, uq) * fac)
                r = np.einsum("qz,cqyx->czyx", S, r)
                r = np.einsum("qy,czqx->czyx", S, r)
                u = np.einsum("qx,czyq->czyx", S, r)
            for kind in np.unique(kinds):
                kind = int(kind)
                if kind == 0:
                    continue
                cells = np.nonzero(kinds == kind)[0]
                v = u[cells]
                for d in range(3):
                    sel = hn_selection(kind, d, lay.degree)
                    if not sel.any():
                        continue
                    s = 1 - ((kind >> d) & 1)
                    Wm = W[s].T if transpose else W[s]
                    vm = np.moveaxis(v, 3 - d, -1)
                    new = vm @ Wm.T
                    vm[:, sel, :] = new[:, sel, :]
                u[cells] = v
        np.add.at(dst, idx.ravel(), u.ravel())
    return dst


def vmult_fast(lay, src, apply_constraints=True, dst=None, chunk=4096):
    """dst += A src with the fast algorithm (accumulating like
    benchmark_03.h:237-241)."""
    n = lay.degree + 1
    if dst is None:
        dst = np.zeros_like(src)
    for c0 in range(0, lay.n_cells, chunk):
        c1 = min(lay.n_cells, c0 + chunk)
        idx = lay.dof_indices[c0:c1]
        u = src[idx].reshape(-1, n, n, n)
        if apply_constraints:
            hn_apply(u, lay.kinds[c0:c1], lay.degree, transpose=False)
        r = cell_laplace(u, lay.h[c0:c1], lay.degree)
        if apply_constraints:
            hn_apply(r, lay.kinds[c0:c1], lay.degree, transpose=True)
        np.add.at(dst, idx.ravel(), r.ravel())
    return dst


# --------------------------------------------------------------------------
# O1: general-purpose operator with explicit constraints from geometry
# --------------------------------------------------------------------------
def reference_element_matrix(degree):
    """K_ref[i][j] = int_[0,1]^3 grad phi_i . grad phi_j with QGauss(k+1),
    built from the nodal derivative matrix G directly (no collocation)."""
    sd = fe1d.shape_data(degree)
    S, G, w = (sd.longdouble(x) for x in ("S", "G", "qw"))
    M1 = (S * w[:, None]).T @ S
    K1 = (G * w[:, None]).T @ G
    K = (np.kron(M1, np.kron(M1, K1)) + np.kron(M1, np.kron(K1, M1)) + np.kron(K1, np.kron(M1, M1)))
    return K  # longdouble, index = ax + n*(ay + n*az)


class GeneralOperator:
    def __init__(self, tree, lay):
        self.lay = lay
        k = lay.degree
        n = k + 1
        nd = lay.n_dofs
        sd = fe1d.shape_data(k)
        cells = [tuple(c) for c in lay.cells.tolist()]
        where = {c: ci for ci, c in enumerate(cells)}
        nodes = tree.has_children
        raw = lay.raw_indices.astype(np.int64)
        Kref = reference_element_matrix(k).astype(np.float64)
        rows = np.repeat(raw, n ** 3, axis=1).ravel()
        cols = np.tile(raw, (1, n ** 3)).ravel()
        vals = (lay.h[:, None] * Kref.ravel()[None, :]).ravel()
        self.K = sp.csr_matrix((vals, (rows, cols)), shape=(nd, nd))
        # constraints from geometry: a raw DoF of a cell that sits on a coarser
        # leaf without being one of that leaf's DoFs is hanging
        x = sd.longdouble("nodes")
        hanging = {}
        offs = [(dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1) if (dx, dy, dz) != (0, 0, 0)]
        for ci, (l, i, j, kk) in enumerate(cells):
            if l == 0:
                continue
            nn = 1 << l
            pos = (i, j, kk)
            for off in offs:
                p = tuple(pos[d] + off[d] for d in range(3))
                if not all(0 <= p[d] < nn for d in range(3)):
                    continue
                if (l,) + p in nodes:
                    continue
                # walk up to the covering leaf
                ll, q = l, p
                while (ll,) + q not in nodes:
                    ll -= 1
                    q = tuple(t >> 1 for t in q)
                assert not nodes[(ll,) + q]
                cn = where[(ll,) + q]
                coarse_set = set(raw[cn].tolist())
                ratio = 1 << (l - ll)
                # local nodes of the fine cell lying on the shared boundary part
                rng = [([0] if off[d] < 0 else [k] if off[d] > 0 else range(n)) for d in range(3)]
                for az in rng[2]:
                    for ay in rng[1]:
                        for ax in rng[0]:
                            g = int(raw[ci, ax + n * (ay + n * az)])
                            if g in coarse_set or g in hanging:
                                continue
                            a = (ax, ay, az)
                            # reference coordinates inside the coarse leaf (exact up to longdouble)
                            xi = [(np.longdouble(pos[d]) + x[a[d]]) / ratio - q[d] for d in range(3)]
                            V = [fe1d.lagrange(x, [xi[d]])[0][0] for d in range(3)]
                            wgt = np.einsum("k,j,i->kji", V[2], V[1], V[0]).ravel().astype(np.float64)
                            nz = np.abs(wgt) > 1e-15
                            hanging[g] = (raw[cn][nz], wgt[nz])
        self.hanging = hanging
        r, c, v = [], [], []
        for g in range(nd):
            if g not in hanging:
                r.append(g), c.append(g), v.append(1.0)
        for g, (cc, ww) in hanging.items():
            assert not any(int(t) in hanging for t in cc), "constraint chain"
            r.extend([g] * len(cc)), c.extend(cc.tolist()), v.extend(ww.tolist())
        self.C = sp.csr_matrix((v, (r, c)), shape=(nd, nd))
        self.is_hanging = np.zeros(nd, dtype=bool)
        self.is_hanging[list(hanging)] = True

    def vmult(self, src):
        return self.C.T @ (self.K @ (self.C @ src))

    def vmult_unconstrained(self, src):
        return self.K @ src
