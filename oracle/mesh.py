"""Single-tree octree meshes on hyper_cube(-1,1)^3 -- oracle restatement
(test infrastructure, see oracle/__init__.py) of the reference's generators:

  create_step               /root/reference/benchmark.h:7-34
  create_quadrant           /root/reference/benchmark.h:38-69  (benchmark_03.h:26-58)
  create_quadrant_flexible  /root/reference/benchmark.h:73-96
  create_annulus            /root/reference/benchmark.h:100-144 (benchmark_03.h:62-104)

deal.II semantics restated here (deal.II itself is absent):
 * ``execute_coarsening_and_refinement`` first closes the refine flags under
   the 2:1 rule -- across faces and, in 3D, across edges for a serial
   ``Triangulation`` (flavour "serial"); p4est's full balance additionally
   across corners (flavour "p4est", parallel::distributed::Triangulation,
   benchmark_03.h:397);
 * children are then created level by level in cell-index order and appended
   to the next level, child c = cx + 2 cy + 4 cz;
 * active cells are iterated by (level, index).
"""
from __future__ import annotations

import itertools
import math

import numpy as np

_FACE = [o for o in itertools.product((-1, 0, 1), repeat=3) if sum(map(abs, o)) == 1]
_EDGE = [o for o in itertools.product((-1, 0, 1), repeat=3) if sum(map(abs, o)) == 2]
_CORNER = [o for o in itertools.product((-1, 0, 1), repeat=3) if sum(map(abs, o)) == 3]


class Octree:
    """Tree nodes are keyed (level, ix, iy, iz); ``levels[l]`` lists the nodes
    of level l in creation (= deal.II index) order."""

    def __init__(self, flavour: str = "serial"):
        assert flavour in ("serial", "p4est")
        self.flavour = flavour
        self.offsets = _FACE + _EDGE + (_CORNER if flavour == "p4est" else [])
        self.levels = [[(0, 0, 0, 0)]]
        self.has_children = {(0, 0, 0, 0): False}

    # -- queries -----------------------------------------------------------
    def active_cells(self):
        """Active cells in deal.II iteration order (level, index)."""
        return [c for lev in self.levels for c in lev if not self.has_children[c]]

    @property
    def n_levels(self):
        return len(self.levels)

    @staticmethod
    def center(c):
        l, i, j, k = c
        h = 2.0 / (1 << l)
        return (-1.0 + (i + 0.5) * h, -1.0 + (j + 0.5) * h, -1.0 + (k + 0.5) * h)

    # -- refinement --------------------------------------------------------
    def refine(self, flagged):
        flagged = set(flagged)
        queue = list(flagged)
        nodes = self.has_children
        while queue:
            l, i, j, k = queue.pop()
            if l == 0:
                continue
            n = 1 << l
            for dx, dy, dz in self.offsets:
                ni, nj, nk = i + dx, j + dy, k + dz
                if not (0 <= ni < n and 0 <= nj < n and 0 <= nk < n):
                    continue
                if (l, ni, nj, nk) in nodes:
                    continue
                p = (l - 1, ni >> 1, nj >> 1, nk >> 1)
                assert p in nodes and not nodes[p], "mesh was not 2:1 balanced"
                if p not in flagged:
                    flagged.add(p)
                    queue.append(p)
        for l in range(len(self.levels)):
            for c in list(self.levels[l]):
                if c in flagged:
                    assert not nodes[c]
                    nodes[c] = True
                    if l + 1 == len(self.levels):
                        self.levels.append([])
                    _, i, j, k = c
                    for ch in range(8):
                        cc = (l + 1, 2 * i + (ch & 1), 2 * j + ((ch >> 1) & 1), 2 * k + ((ch >> 2) & 1))
                        nodes[cc] = False
                        self.levels[l + 1].append(cc)

    def refine_global(self, times=1):
        for _ in range(times):
            self.refine(self.active_cells())

    def refine_if(self, pred):
        self.refine([c for c in self.active_cells() if pred(self.center(c))])


def create_quadrant(n_refinements, flavour="serial"):
    t = Octree(flavour)
    if n_refinements == 0:
        return t
    t.refine_global(1)
    for _ in range(1, n_refinements):
        t.refine_if(lambda c: all(x <= 0.0 for x in c))
    assert t.n_levels - 1 == n_refinements
    return t


def create_step(n_refinements, flavour="serial"):
    t = Octree(flavour)
    if n_refinements == 0:
        return t
    t.refine_global(1)
    for _ in range(1, n_refinements):
        t.refine_if(lambda c: c[0] <= 0.0)
    assert t.n_levels - 1 == n_refinements
    return t


def create_quadrant_flexible(n_ref_global, n_ref_local=1, flavour="serial"):
    t = Octree(flavour)
    t.refine_global(n_ref_global)
    for _ in range(n_ref_local):
        t.refine_if(lambda c: all(x <= 0.0 for x in c))
    return t


def _norm(c):
    return math.sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2])


def create_annulus(n_refinements, flavour="serial"):
    t = Octree(flavour)
    if n_refinements == 0:
        return t
    for _ in range(n_refinements - 3):
        t.refine_global(1)
    if n_refinements >= 1:
        t.refine_if(lambda c: _norm(c) < 0.55)
    if n_refinements >= 2:
        t.refine_if(lambda c: 0.3 <= _norm(c) <= 0.43)
    if n_refinements >= 3:
        t.refine_if(lambda c: 0.335 <= _norm(c) <= 0.39)
    return t


GENERATORS = {
    "quadrant": create_quadrant,
    "annulus": create_annulus,
    "step": create_step,
    "quadrant_flexible": create_quadrant_flexible,
}


def create(geometry, n_refinements, flavour="serial"):
    if geometry not in GENERATORS:
        raise ValueError("Unknown geometry type!")  # benchmark_01.h:217
    return GENERATORS[geometry](n_refinements, flavour=flavour)


def cells_array(tree):
    """int32[n_cells,4] (level, ix, iy, iz) of the active cells in storage order."""
    return np.array(tree.active_cells(), dtype=np.int32).reshape(-1, 4)
