"""FE_Q(k) DoF enumeration, hanging-node masks and coarse-index substitution
-- oracle restatement (test infrastructure, see oracle/__init__.py).

What is restated (deal.II is absent):
 * ``DoFHandler::distribute_dofs(FE_Q(k))`` (reference call sites
   benchmark_01.h:247, benchmark_03.h:438-439): active cells are walked in
   storage order and every not-yet-numbered object receives consecutive
   indices in the order vertices 0-7, lines 0-11, quads 0-5, hex interior;
   hanging vertices / lines / quads own DoFs too (``n_dofs()`` counts them).
 * ``ConstraintKinds`` / ``compress`` / ``decompress`` (reference use:
   benchmark_00_likwid.cc:41-48, benchmark_01.h:318-347).
 * ``HangingNodes::setup_constraints``: detection of coarser face / edge
   neighbours (spec twin: constraint_helper.h:89-125) and replacement of the
   cell's face / edge DoF indices by the coarse neighbour's.
"""
from __future__ import annotations

import numpy as np

from . import fe1d

# deal.II GeometryInfo<3>: vertex v = vx + 2 vy + 4 vz
# lines: (direction, fixed coordinates of the two transversal directions)
#   0-3 on z=0: {x=0,along y},{x=1,along y},{y=0,along x},{y=1,along x}; 4-7 same on z=1;
#   8-11 along z at (x,y) = (0,0),(1,0),(0,1),(1,1)   (cf. constraint_helper.h:21-32)
LINES = [
    (1, {0: 0, 2: 0}), (1, {0: 1, 2: 0}), (0, {1: 0, 2: 0}), (0, {1: 1, 2: 0}),
    (1, {0: 0, 2: 1}), (1, {0: 1, 2: 1}), (0, {1: 0, 2: 1}), (0, {1: 1, 2: 1}),
    (2, {0: 0, 1: 0}), (2, {0: 1, 1: 0}), (2, {0: 0, 1: 1}), (2, {0: 1, 1: 1}),
]
# quads: (normal, side, fast tangential dir, slow tangential dir); face-local
# coordinate systems are (y,z), (z,x), (x,y) for normals x, y, z.
QUADS = [(0, 0, 1, 2), (0, 1, 1, 2), (1, 0, 2, 0), (1, 1, 2, 0), (2, 0, 0, 1), (2, 1, 0, 1)]


def compress(kind: int) -> int:
    """ConstraintKinds (uint16: subcell bits 0-2, face 3-5, edge 6-8) -> uint8."""
    subcell = kind & 7
    face = (kind >> 3) & 7
    edge = (kind >> 6) & 7
    return subcell + ((face > 0) << 3) + ((edge > 0) << 4) + (max(face, edge) << 5)


def decompress(byte: int) -> int:
    subcell = byte & 7
    flag0 = (byte >> 3) & 3
    flag1 = (byte >> 5) & 7
    return subcell + ((flag1 if (flag0 & 1) else 0) << 3) + ((flag1 if (flag0 & 2) else 0) << 6)


def check(kind: int) -> bool:
    """Valid 3D kinds: unconstrained; faces only; edges only; one face plus the
    edge of the same letter."""
    if kind == 0:
        return True
    if kind >> 9:
        return False
    face = (kind >> 3) & 7
    edge = (kind >> 6) & 7
    if face == 0 and edge == 0:
        return False  # subcell bits without a constraint
    if face and edge:
        return face == edge and face in (1, 2, 4)
    return True


def valid_kinds():
    return [k for k in range(512) if k != 0 and check(k)]


class DoFLayout:
    """Result of the setup on one mesh + degree.

    cells            int32[n_cells,4]   (level, ix, iy, iz), storage order
    n_dofs           like dof_handler.n_dofs() (hanging DoFs included)
    raw_indices      uint32[n_cells,n^3] lexicographic (x fastest), before substitution
    dof_indices      uint32[n_cells,n^3] after coarse-index substitution
    kinds            uint16[n_cells]    ConstraintKinds
    masks            uint8[n_cells]     compress(kinds)
    h                float64[n_cells]   edge length of the (Cartesian) cell
    support_points   float64[n_dofs,3]
    """


def lex(a, n):
    return a[0] + n * (a[1] + n * a[2])


def distribute_dofs(tree, degree: int, cell_order=None):
    """Enumerate FE_Q(degree) DoFs over the active cells of ``tree``.

    Returns raw (unsubstituted) lexicographic cell index arrays and n_dofs.
    ``cell_order`` optionally gives the walk order (used for multi-rank
    numbering: ranks in ascending order, each rank's cells in storage order)."""
    k = degree
    n = k + 1
    cells = tree.active_cells() if cell_order is None else cell_order
    lmax = tree.n_levels - 1
    vtx, line, quad = {}, {}, {}
    nxt = 0
    out = np.empty((len(cells), n ** 3), dtype=np.int64)
    km1 = k - 1
    for ci, (l, i, j, kk) in enumerate(cells):
        s = 1 << (lmax - l)
        o = (i * s, j * s, kk * s)
        idx = out[ci]
        # vertices
        for v in range(8):
            b = (v & 1, (v >> 1) & 1, (v >> 2) & 1)
            key = (o[0] + b[0] * s, o[1] + b[1] * s, o[2] + b[2] * s)
            g = vtx.get(key)
            if g is None:
                g = vtx[key] = nxt
                nxt += 1
            idx[lex((b[0] * k, b[1] * k, b[2] * k), n)] = g
        if km1 > 0:
            for d, fixed in LINES:
                org = [o[0], o[1], o[2]]
                for t, side in fixed.items():
                    org[t] += side * s
                key = (l, d, org[0], org[1], org[2])
                g = line.get(key)
                if g is None:
                    g = line[key] = nxt
                    nxt += km1
                a = [0, 0, 0]
                for t, side in fixed.items():
                    a[t] = side * k
                for m in range(km1):
                    a[d] = m + 1
                    idx[lex(a, n)] = g + m
            for nd, side, fast, slow in QUADS:
                org = [o[0], o[1], o[2]]
                org[nd] += side * s
                key = (l, nd, org[0], org[1], org[2])
                g = quad.get(key)
                if g is None:
                    g = quad[key] = nxt
                    nxt += km1 * km1
                a = [0, 0, 0]
                a[nd] = side * k
                for ms in range(km1):
                    a[slow] = ms + 1
                    for mf in range(km1):
                        a[fast] = mf + 1
                        idx[lex(a, n)] = g + mf + km1 * ms
            for az in range(km1):
                for ay in range(km1):
                    for ax in range(km1):
                        idx[lex((ax + 1, ay + 1, az + 1), n)] = nxt
                        nxt += 1
    return out, nxt


def constraint_kinds(tree, cells=None):
    """ConstraintKinds of every active cell (Appendix B items 1-2; equivalent to
    Helper::is_constrained, constraint_helper.h:89-125, under 2:1 balance)."""
    cells = tree.active_cells() if cells is None else cells
    nodes = tree.has_children
    kinds = np.zeros(len(cells), dtype=np.uint16)
    for ci, (l, i, j, kk) in enumerate(cells):
        if l == 0:
            continue
        n = 1 << l
        pos = (i, j, kk)
        b = (i & 1, j & 1, kk & 1)
        out = tuple(pos[d] + 2 * b[d] - 1 for d in range(3))  # parent's outer side

        def coarser(p):
            if not all(0 <= p[d] < n for d in range(3)):
                return False
            return (l, p[0], p[1], p[2]) not in nodes

        face = [False] * 3
        for d in range(3):
            p = list(pos)
            p[d] = out[d]
            face[d] = coarser(p)
        edge = [False] * 3
        for d in range(3):
            a, bb = [t for t in range(3) if t != d]
            if face[a] or face[bb]:
                continue
            p = list(pos)
            p[a] = out[a]
            p[bb] = out[bb]
            edge[d] = coarser(p)
        if any(face) or any(edge):
            kind = 0
            for d in range(3):
                kind |= (1 - b[d]) << d
                kind |= int(face[d]) << (3 + d)
                kind |= int(edge[d]) << (6 + d)
            kinds[ci] = kind
    return kinds


def substitute(tree, cells, raw, kinds, degree):
    """Replace constrained face / edge slots by the coarse neighbour's DoFs
    (Appendix B item 3)."""
    k = degree
    n = k + 1
    where = {c: ci for ci, c in enumerate(cells)}
    out = raw.copy()
    for ci, (l, i, j, kk) in enumerate(cells):
        kind = int(kinds[ci])
        if kind == 0:
            continue
        pos = (i, j, kk)
        b = (i & 1, j & 1, kk & 1)
        for d in range(3):
            if kind & (1 << (3 + d)):
                p = [pos[0] >> 1, pos[1] >> 1, pos[2] >> 1]
                p[d] += 2 * b[d] - 1
                cn = where[(l - 1, p[0], p[1], p[2])]
                t0, t1 = [t for t in range(3) if t != d]
                a = [0, 0, 0]
                an = [0, 0, 0]
                a[d] = b[d] * k
                an[d] = (1 - b[d]) * k
                for m1 in range(n):
                    for m0 in range(n):
                        a[t0] = an[t0] = m0
                        a[t1] = an[t1] = m1
                        out[ci, lex(a, n)] = raw[cn, lex(an, n)]
        for d in range(3):
            if kind & (1 << (6 + d)):
                t0, t1 = [t for t in range(3) if t != d]
                p = [pos[0] >> 1, pos[1] >> 1, pos[2] >> 1]
                p[t0] += 2 * b[t0] - 1
                p[t1] += 2 * b[t1] - 1
                cn = where[(l - 1, p[0], p[1], p[2])]
                a = [0, 0, 0]
                an = [0, 0, 0]
                a[t0], a[t1] = b[t0] * k, b[t1] * k
                an[t0], an[t1] = (1 - b[t0]) * k, (1 - b[t1]) * k
                for m in range(n):
                    a[d] = an[d] = m
                    out[ci, lex(a, n)] = raw[cn, lex(an, n)]
    return out


def support_points(cells, raw, n_dofs, degree):
    sd = fe1d.shape_data(degree)
    x = sd.nodes
    n = degree + 1
    pts = np.full((n_dofs, 3), np.nan)
    cells = np.asarray(cells)
    h = 2.0 / (1 << cells[:, 0]).astype(np.float64)
    ax = np.arange(n ** 3) % n
    ay = (np.arange(n ** 3) // n) % n
    az = np.arange(n ** 3) // (n * n)
    for d, a in enumerate((ax, ay, az)):
        coord = -1.0 + (cells[:, 1 + d, None] + x[a][None, :]) * h[:, None]
        pts[raw.ravel(), d] = coord.ravel()
    return pts


def setup(tree, degree: int) -> DoFLayout:
    cells = tree.active_cells()
    raw, n_dofs = distribute_dofs(tree, degree)
    kinds = constraint_kinds(tree, cells)
    sub = substitute(tree, cells, raw, kinds, degree)
    lay = DoFLayout()
    lay.degree = degree
    lay.cells = np.array(cells, dtype=np.int32).reshape(-1, 4)
    lay.n_cells = len(cells)
    lay.n_dofs = n_dofs
    lay.raw_indices = raw.astype(np.uint32)
    lay.dof_indices = sub.astype(np.uint32)
    lay.kinds = kinds
    lay.masks = np.array([compress(int(x)) for x in kinds], dtype=np.uint8)
    lay.h = 2.0 / (1 << lay.cells[:, 0]).astype(np.float64)
    lay.support_points = support_points(cells, raw, n_dofs, degree)
    return lay
