"""Restatement of the reference's ``Helper<dim>`` (test infrastructure, see
oracle/__init__.py): /root/reference/constraint_helper.h:8-125 -- the
reference's own definition of "cell with hanging nodes", used for the
n_cells_hn column of benchmark_03 (benchmark_03.h:415-432) and for the
partition weights of benchmark_02 (benchmark_02.cc:15-37).

This is real reference code (not deal.II), so it pins the constraint
detection of oracle/dofs.py and of the C++ setup: ``is_constrained(cell)`` must
hold exactly for the cells whose ConstraintKinds is non-zero.
"""
from __future__ import annotations

# constraint_helper.h:21-32: which two children of a refined cell share (a half of) its line l
LINE_TO_CHILDREN = [(0, 2), (1, 3), (0, 1), (2, 3), (4, 6), (5, 7), (4, 5), (6, 7), (0, 4), (1, 5), (2, 6), (3, 7)]

# deal.II GeometryInfo<3> lines: (direction, {transversal dim: side})
LINES = [
    (1, {0: 0, 2: 0}), (1, {0: 1, 2: 0}), (0, {1: 0, 2: 0}), (0, {1: 1, 2: 0}),
    (1, {0: 0, 2: 1}), (1, {0: 1, 2: 1}), (0, {1: 0, 2: 1}), (0, {1: 1, 2: 1}),
    (2, {0: 0, 1: 0}), (2, {0: 1, 1: 0}), (2, {0: 0, 1: 1}), (2, {0: 1, 1: 1}),
]


def _line_key(cell, ln):
    """Identity of line `ln` of tree node `cell` = (level, direction, origin in level units)."""
    l, i, j, k = cell
    d, fixed = LINES[ln]
    org = [i, j, k]
    for t, side in fixed.items():
        org[t] += side
    return (l, d, org[0], org[1], org[2])


def _child(cell, c):
    l, i, j, k = cell
    return (l + 1, 2 * i + (c & 1), 2 * j + ((c >> 1) & 1), 2 * k + ((c >> 2) & 1))


class Helper:
    def __init__(self, tree):
        self.tree = tree
        nodes = tree.has_children
        line_to_cells, line_to_inactive = {}, {}
        # constraint_helper.h:40-56: add active and inactive cells to their lines
        for cell, refined in nodes.items():
            for ln in range(12):
                (line_to_inactive if refined else line_to_cells).setdefault(_line_key(cell, ln), []).append((cell, ln))
        # constraint_helper.h:62-84: the children of a line that has active and inactive cells around it
        # inherit the active cells of the parent line
        for key, active in list(line_to_cells.items()):
            inactive = line_to_inactive.get(key)
            if not active or not inactive:
                continue
            inactive_cell, neighbor_line = inactive[0]
            for c in range(2):
                child = _child(inactive_cell, LINE_TO_CHILDREN[neighbor_line][c])
                line_to_cells.setdefault(_line_key(child, neighbor_line), []).extend(active)
        self.line_to_cells = line_to_cells

    def _neighbor_level(self, cell, d, side):
        """Level of cell->neighbor(face) for an active cell: the same-level neighbour if it exists
        (active or refined), otherwise the coarser active cell covering it; None at the boundary."""
        l, i, j, k = cell
        p = [i, j, k]
        p[d] += 1 if side else -1
        if not 0 <= p[d] < (1 << l):
            return None
        ll, q = l, tuple(p)
        while (ll,) + q not in self.tree.has_children:
            ll -= 1
            q = tuple(t >> 1 for t in q)
        return ll

    def is_face_constrained(self, cell):  # constraint_helper.h:97-108
        for d in range(3):
            for side in (0, 1):
                nl = self._neighbor_level(cell, d, side)
                if nl is not None and cell[0] > nl:
                    return True
        return False

    def is_edge_constrained(self, cell):  # constraint_helper.h:110-123
        for ln in range(12):
            for other, _ in self.line_to_cells.get(_line_key(cell, ln), []):
                if cell[0] > other[0]:
                    return True
        return False

    def is_constrained(self, cell):  # constraint_helper.h:89-95
        return self.is_face_constrained(cell) or self.is_edge_constrained(cell)
