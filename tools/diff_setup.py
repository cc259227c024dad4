#!/usr/bin/env python
"""Bit-for-bit comparison of a setup dump written by the deal.II-side shim (integration/mfhn_dealii.h compiled with
-DMFHN_DEALII_DUMP_SETUP='"setup.bin"': MatrixFree::get_dof_info().dof_indices, hanging_node_constraint_masks, cell
extents, in MatrixFree's own cell order) with this engine's setup layer on the same mesh and degree.

    python tools/diff_setup.py setup.bin annulus 5 4 [serial|p4est]
    python tools/diff_setup.py --self-test                     # writes a dump from this engine's arrays and compares it

MatrixFree orders its cells as it likes, so cells are matched through their sorted DoF index rows; reported are the number
of cells without a partner, and for matched cells the mismatches of the lexicographic index order and of the mask byte.
Exit code 0 = identical.  Needs no GPU."""
import importlib
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def read_dump(path):
    with open(path, "rb") as f:
        n_cells, n3 = struct.unpack("<qq", f.read(16))
        idx = np.frombuffer(f.read(4 * n_cells * n3), dtype="<u4").reshape(n_cells, n3)
        masks = np.frombuffer(f.read(n_cells), dtype=np.uint8)
        h = np.frombuffer(f.read(8 * n_cells), dtype="<f8")
    return idx, masks, h


def write_dump(path, idx, masks, h):
    with open(path, "wb") as f:
        f.write(struct.pack("<qq", idx.shape[0], idx.shape[1]))
        f.write(np.ascontiguousarray(idx, dtype="<u4").tobytes())
        f.write(np.ascontiguousarray(masks, dtype=np.uint8).tobytes())
        f.write(np.ascontiguousarray(h, dtype="<f8").tobytes())


def diff(idx_a, masks_a, h_a, idx_b, masks_b, h_b):
    """a: dump, b: this engine.  Returns (cells without partner, index-order mismatches, mask mismatches, h mismatches)."""
    if idx_a.shape != idx_b.shape:
        print(f"shapes differ: dump {idx_a.shape}, engine {idx_b.shape}")
        return max(idx_a.shape[0], idx_b.shape[0]), 0, 0, 0
    key = lambda idx: [bytes(np.sort(r).astype("<u4").tobytes()) for r in idx]  # noqa: E731
    where = {}
    for c, k in enumerate(key(idx_b)):
        where.setdefault(k, []).append(c)
    lonely = order = mask = geom = 0
    for c, k in enumerate(key(idx_a)):
        cand = where.get(k)
        if not cand:
            lonely += 1
            continue
        p = cand.pop()
        order += int(not np.array_equal(idx_a[c], idx_b[p]))
        mask += int(masks_a[c] != masks_b[p])
        geom += int(h_a[c] != h_b[p])
    return lonely, order, mask, geom


def engine_arrays(geo, L, k, flavour):
    mfhn = importlib.import_module("dealii-matrixfree-hanging-nodes_b200")
    tria = mfhn.Triangulation(geo, L, flavour)
    mf = mfhn.MatrixFree(mfhn.DoFHandler(tria, k))
    return mf.dof_indices, mf.masks, mf.h


def main():
    if sys.argv[1:] == ["--self-test"]:
        import tempfile

        idx, masks, h = engine_arrays("annulus", 4, 3, "p4est")
        perm = np.random.default_rng(0).permutation(idx.shape[0])  # another cell order, as MatrixFree would choose
        with tempfile.TemporaryDirectory() as d:
            write_dump(os.path.join(d, "s.bin"), idx[perm], masks[perm], h[perm])
            a = read_dump(os.path.join(d, "s.bin"))
        assert diff(*a, idx, masks, h) == (0, 0, 0, 0)
        bad = a[1].copy()
        bad[3] ^= 0x20
        assert diff(a[0], bad, a[2], idx, masks, h) == (0, 0, 1, 0)
        print("self-test passed")
        return 0
    path, geo, L, k = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    flavour = sys.argv[5] if len(sys.argv) > 5 else "p4est"
    a = read_dump(path)
    res = diff(*a, *engine_arrays(geo, L, k, flavour))
    print(f"cells without partner: {res[0]}, index-order mismatches: {res[1]}, mask mismatches: {res[2]}, extent mismatches: {res[3]}")
    return int(any(res))


if __name__ == "__main__":
    sys.exit(main())
