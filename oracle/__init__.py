"""CPU oracle for the 3D Laplace vmult with hanging-node constraints.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import or execute it, and
there only as the checker / reported CPU baseline.

PARITY UNPINNED: the reference repository (/root/reference) holds no tests,
golden vectors or fixtures for this path, and its arithmetic lives in an
un-vendored, un-pinned deal.II fork (branch ``compressed_constraint_kind_use``,
README.md:16-59) that cannot be built here.  The oracle therefore restates the
published algorithm from the reference's call sites and is pinned by first
principles instead: two independently coded operators (O1 general-purpose
constraints, O2 fast hanging-node algorithm) must agree, and both must
reproduce analytic known answers (see tests/).
"""
