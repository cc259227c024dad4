"""The oracle pinned by first principles (the reference holds no tests, golden
vectors or fixtures for this path -- SURVEY.md 4, 8c): 1D FE identities, mesh
statistics, mask encoding, two independently coded operators, analytic known
answers."""
import numpy as np
import pytest

from oracle import cpu, dofs, fe1d, mesh, operators


@pytest.mark.parametrize("k", range(1, 9))
def test_fe1d_identities(k):
    sd = fe1d.shape_data(k)
    n = k + 1
    assert np.allclose(sd.S.sum(axis=1), 1, atol=1e-14)  # partition of unity
    assert np.allclose(sd.G.sum(axis=1), 0, atol=1e-12)
    assert abs(sd.qw.sum() - 1) < 1e-15
    for p in range(k + 1):  # exact interpolation / differentiation of x^p
        f, df = sd.nodes ** p, p * sd.qpts ** max(p - 1, 0) * (p > 0)
        assert np.allclose(sd.S @ f, sd.qpts ** p, atol=1e-13)
        assert np.allclose(sd.G @ f, df, atol=1e-11)
        assert np.allclose(sd.Dc @ (sd.S @ f), df, atol=1e-11)  # collocation gradient = nodal gradient
    # W_1[i][j] = W_0[k-i][k-j] and polynomial reproduction on the sub-intervals
    assert np.allclose(sd.W[1], sd.W[0][::-1, ::-1], atol=1e-15)
    for s in (0, 1):
        for p in range(k + 1):
            assert np.allclose(sd.W[s] @ sd.nodes ** p, ((sd.nodes + s) / 2) ** p, atol=1e-13)
    # Gauss quadrature with k+1 points integrates the 1D mass / stiffness products exactly
    M = (sd.S * sd.qw[:, None]).T @ sd.S
    assert abs(M.sum() - 1) < 1e-14
    assert sd.S.shape == (n, n)


# cells, cells with hanging nodes (Helper::is_constrained, constraint_helper.h:89-125): SURVEY.md 6
MESH_TABLE = [("quadrant", 2, "serial", 15, None), ("quadrant", 5, "serial", 4705, 1082), ("quadrant", 6, "serial", 34896, 4257),
              ("quadrant", 6, "p4est", 34903, 4258), ("annulus", 6, "serial", 6616, 5328), ("annulus", 4, "serial", 8, 0)]


@pytest.mark.parametrize("geo,L,flavour,n_cells,n_hn", MESH_TABLE)
def test_mesh_statistics(geo, L, flavour, n_cells, n_hn):
    t = mesh.create(geo, L, flavour)
    cells = t.active_cells()
    assert len(cells) == n_cells
    if n_hn is not None:
        assert int((dofs.constraint_kinds(t, cells) != 0).sum()) == n_hn
    if geo == "quadrant":
        assert t.n_levels - 1 == L  # benchmark.h:68


def test_unknown_geometry_raises():
    with pytest.raises(ValueError, match="Unknown geometry type"):  # benchmark_01.h:217
        mesh.create("torus", 3)


@pytest.mark.parametrize("k,n_dofs", [(1, 5696), (2, 42411), (3, 138034), (4, 320795)])
def test_dof_counts_quadrant5(k, n_dofs):
    raw, nd = dofs.distribute_dofs(mesh.create("quadrant", 5), k)
    assert nd == n_dofs  # dof_handler.n_dofs() counts hanging DoFs
    assert sorted(np.unique(raw)) == list(range(nd))
    assert dofs.distribute_dofs(mesh.create("quadrant", 2), 1)[1] == 46


def test_constraint_kind_encoding():
    valid = dofs.valid_kinds()
    assert len(valid) == 136  # 8 subcells x (7 face sets + 7 edge sets + 3 face+edge pairs)
    compressed = [dofs.compress(k) for k in valid]
    assert len(set(compressed)) == 136 and 0 not in compressed and max(compressed) < 256
    for k in valid:
        assert dofs.decompress(dofs.compress(k)) == k
    assert dofs.compress(0) == 0 and dofs.decompress(0) == 0
    assert sum(dofs.check(k) for k in range(512)) == 137
    # benchmark_00_likwid.cc:41-48: subcell 1, all three faces
    kind = 1 + (7 << 3) + (0 << 6)
    assert dofs.check(kind) and dofs.compress(kind) == 1 + (1 << 3) + (7 << 5)


def test_masks_on_meshes_are_valid():
    for geo, L in (("quadrant", 4), ("annulus", 5)):
        t = mesh.create(geo, L)
        kinds = dofs.constraint_kinds(t)
        assert all(dofs.check(int(k)) for k in kinds)
    assert len(np.unique(kinds)) == 105  # 104 constrained kinds occur on annulus L=5


@pytest.mark.parametrize("geo,L,ks", [("quadrant", 3, (1, 2, 3, 4, 5)), ("annulus", 5, (1, 2, 3)), ("quadrant", 2, (6, 7, 8))])
def test_fast_algorithm_equals_general_purpose_operator(geo, L, ks):
    """O1 (explicit constraints from geometry, no masks) == O2 (fast algorithm):
    the cross-check the reference sets up but never evaluates (benchmark_01.h:286-293)."""
    t = mesh.create(geo, L)
    rng = np.random.default_rng(12345)
    for k in ks:
        lay = dofs.setup(t, k)
        O1 = operators.GeneralOperator(t, lay)
        x = rng.uniform(-1, 1, lay.n_dofs)
        y1, y2 = O1.vmult(x), operators.vmult_fast(lay, x)
        assert np.abs(y1 - y2).max() / np.abs(y1).max() < 1e-13
        # the test discriminates: without interpolation the result is wrong by O(1)
        y3 = operators.vmult_fast(lay, x, apply_constraints=False)
        assert np.abs(y1 - y3).max() / np.abs(y1).max() > 1e-2
        # hanging entries are never written
        assert np.abs(y2[O1.is_hanging]).max() == 0
        live = ~O1.is_hanging
        assert (np.abs(y2[live]) > 0).all()


@pytest.mark.parametrize("k", [1, 2, 3, 4])
def test_known_answers(k):
    """A 1 = 0 (src == 1 is what benchmark_01.h:510-511 times), u^T A u = int |grad p|^2
    for polynomials of degree <= k on hanging-node meshes, symmetry."""
    import sympy as sp

    x, y, z = sp.symbols("x y z")
    p_sym = x + 2 * y - z + x * y * z if k == 1 else x * x + y * z + sp.Rational(1, 2) * x * y * z + x * y * y
    exact = float(sp.integrate(sum(sp.diff(p_sym, v) ** 2 for v in (x, y, z)), (x, -1, 1), (y, -1, 1), (z, -1, 1)))
    f = sp.lambdify((x, y, z), p_sym, "numpy")
    for geo, L in (("quadrant", 3), ("annulus", 5)):
        lay = dofs.setup(mesh.create(geo, L), k)
        pts = lay.support_points
        u = f(pts[:, 0], pts[:, 1], pts[:, 2])
        Au = operators.vmult_fast(lay, u)
        assert abs(u @ Au - exact) / exact < 1e-13
        assert abs(u @ operators.vmult_fast(lay, u, False) - exact) / exact > 1e-2
        assert np.abs(operators.vmult_fast(lay, np.ones(lay.n_dofs))).max() < 1e-13
        rng = np.random.default_rng(12345)
        a, b = rng.uniform(-1, 1, lay.n_dofs), rng.uniform(-1, 1, lay.n_dofs)
        Aa, Ab = operators.vmult_fast(lay, a), operators.vmult_fast(lay, b)
        live = Aa != 0
        assert abs(b[live] @ Aa[live] - a[live] @ Ab[live]) < 1e-11 * abs(b[live] @ Aa[live])
        assert a[live] @ Aa[live] > 0


def test_energy_of_sin_converges():
    """u = sum sin x_d: u^T A u -> 12 (1 + sin(2)/2) = 17.4558... (benchmark_03.h:362-378 vector)."""
    exact = 12 * (1 + np.sin(2.0) / 2)
    lay = dofs.setup(mesh.create("annulus", 5), 4)
    u = np.sin(lay.support_points).sum(axis=1)
    assert abs(u @ operators.vmult_fast(lay, u) - exact) / exact < 1e-6


@pytest.mark.parametrize("k", [1, 2, 3, 5])
def test_hanging_node_kernel_adjoint_and_reproduction(k):
    """<W x, y> = <x, W^T y> for all 136 kinds; polynomial reproduction on constrained faces."""
    rng = np.random.default_rng(3)
    kinds = np.array(dofs.valid_kinds(), dtype=np.uint16)
    n = k + 1
    x, y = rng.uniform(-1, 1, (len(kinds), n, n, n)), rng.uniform(-1, 1, (len(kinds), n, n, n))
    Wx = operators.hn_apply(x.copy(), kinds, k, False)
    Wty = operators.hn_apply(y.copy(), kinds, k, True)
    assert np.allclose((Wx * y).sum(axis=(1, 2, 3)), (x * Wty).sum(axis=(1, 2, 3)), rtol=1e-12, atol=1e-12)
    # pass order is irrelevant
    assert not np.allclose(Wx, x)


@pytest.mark.parametrize("geo,L,k", [("annulus", 5, 1), ("annulus", 5, 4), ("quadrant", 4, 6), ("quadrant", 3, 8)])
def test_c_restatement_matches_numpy(geo, L, k):
    lay = dofs.setup(mesh.create(geo, L), k)
    x = np.random.default_rng(1).uniform(-1, 1, lay.n_dofs)
    for ac in (True, False):
        ref = operators.vmult_fast(lay, x, ac)
        y = cpu.vmult(k, lay.dof_indices, lay.masks, lay.h, x, apply_constraints=ac)
        assert np.abs(y - ref).max() / np.abs(ref).max() < 1e-14
