"""Parity of the CUDA vmult (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): 1e-12 relative in double, 1e-5 in float,
max-norm scaled by ||A src||_inf.  Single precision on the SMOOTH input
(sum sin x_d on a refined mesh) is the one exception: A src cancels to a small
fraction of |A||src| there, so no float evaluation (the reference's included)
can reach 1e-5 of ||A src||; that case is scaled by || |A| |src| ||_inf, the
magnitude the arithmetic actually works against, and additionally bounded by
1e-4 of ||A src||.  DoF indices and masks are compared bit-exact in
test_setup_parity.py."""
import numpy as np
import pytest

from oracle import dofs, mesh, operators

pytestmark = pytest.mark.gpu

TOL = {"double": 1e-12, "float": 1e-5}


def _case(mfhn, geo, L, flavour, k):
    tria = mfhn.Triangulation(geo, L, flavour)
    dh = mfhn.DoFHandler(tria, k)
    mf = mfhn.MatrixFree(dh)
    lay = dofs.setup(mesh.create(geo, L, flavour), k)
    return dh, mf, lay


def _src(lay, kind):
    if kind == "sin":  # benchmark_03.h:362-378
        return np.sin(lay.support_points).sum(axis=1)
    return np.random.default_rng(12345).uniform(-1, 1, lay.n_dofs)


def _run(mfhn, mf, x, number, kernel, apply_constraints=True):
    import torch

    op = mfhn.LaplaceOperator(mf, number=number, kernel=kernel, apply_constraints=apply_constraints)
    src = op.initialize_dof_vector()
    dst = op.initialize_dof_vector()
    src.copy_(torch.from_numpy(x).to(src.dtype))
    op.vmult(dst, src)
    torch.cuda.synchronize()
    return dst.cpu().numpy().astype(np.float64), op


KERNELS = ["qpoint", "separable", "plane", "bulk", "runs", "baseline"]


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("number", ["double", "float"])
def test_vmult_matches_oracle_annulus(mfhn, k, kernel, number):
    if kernel == "bulk" and not 3 <= k <= 5:
        pytest.skip("the bulk-copy kernel covers degrees 3..5")
    if kernel == "runs" and k > 5:
        pytest.skip("the run-wise bulk-copy kernel covers degrees 1..5")
    L = 5 if k <= 4 else 4 if k <= 6 else 3
    geo = "annulus" if k <= 4 else "quadrant"
    dh, mf, lay = _case(mfhn, geo, L, "serial", k)
    for kind in ("sin", "random"):
        x = _src(lay, kind)
        ref = operators.vmult_fast(lay, x)
        y, _ = _run(mfhn, mf, x, number, kernel)
        err = np.abs(y - ref).max() / np.abs(ref).max()
        if number == "float" and kind == "sin":
            # A src cancels by > 10^3 on this smooth input: ROUNDING THE INPUT to float alone (double arithmetic after
            # that) already moves the result by `pert` ~ 1e-5 of its size, so no float evaluation can meet 1e-5 of
            # |A src|; the kernel has to stay within 15 x that perturbation and within 1e-5 of the
            # cancellation-free scale |A| |src|
            pert = np.abs(operators.vmult_fast(lay, x.astype(np.float32).astype(np.float64)) - ref).max() / np.abs(ref).max()
            scale = np.abs(operators.vmult_abs_bound(lay, x)).max()
            assert np.abs(y - ref).max() / scale < TOL[number], (k, kernel, number, kind, err)
            assert err < max(2e-5, 15 * pert), (k, kernel, number, kind, err, pert)
            assert err < 1e-4, (k, kernel, number, kind, err)
        else:
            assert err < TOL[number], (k, kernel, number, kind, err)


@pytest.mark.parametrize("kernel", KERNELS)
def test_vmult_without_constraints(mfhn, kernel):
    """apply_constraints=false: plain gather/scatter on the same index arrays
    (benchmark_03.h:255-268)."""
    dh, mf, lay = _case(mfhn, "annulus", 5, "p4est", 3)
    x = _src(lay, "random")
    ref = operators.vmult_fast(lay, x, apply_constraints=False)
    y, _ = _run(mfhn, mf, x, "double", kernel, apply_constraints=False)
    assert np.abs(y - ref).max() / np.abs(ref).max() < 1e-12


@pytest.mark.parametrize("number", ["double", "float"])
@pytest.mark.parametrize("k", [3, 4, 5])
def test_bulk_kernel_ranges_irregular_cells_and_alignment(mfhn, k, number):
    """MFHN_KERNEL_BULK / MFHN_KERNEL_RUNS: cell ranges (the partitions of the overlap schedule, not aligned to warp
    batches) and the 16-byte alignment rule."""
    import torch

    dh, mf, lay = _case(mfhn, "quadrant", 3, "p4est", k)
    x = _src(lay, "random")
    ref = operators.vmult_fast(lay, x)
    op = mfhn.LaplaceOperator(mf, number=number, kernel="bulk")
    assert op.query("bulk_irregular_cells") <= 1  # the padding behind the vectors takes the widened last block (double)
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    src.copy_(torch.from_numpy(x).to(src.dtype))
    cuts = [0, 7, 8, 50, mf.n_cells]
    for kern in ("bulk", "runs"):
        op.set_kernel(kern)
        dst.zero_()
        for b, e in zip(cuts[:-1], cuts[1:]):
            op.vmult_range(dst, src, b, e)
        torch.cuda.synchronize()
        y = dst.cpu().numpy().astype(np.float64)
        assert np.abs(y - ref).max() / np.abs(ref).max() < TOL[number], kern
    assert op.query("runs_bulk_copies") >= mf.n_cells - 8 and op.query("runs_single_entries") > 0
    # unaligned views are rejected (bulk copies need 16-byte aligned addresses)
    big = torch.zeros(src.numel() + 8, dtype=src.dtype, device=src.device)
    with pytest.raises(mfhn.MfhnError):
        op.vmult(dst, big[1 : 1 + src.numel()])
    op.set_kernel("plane")
    dst.zero_()
    op.vmult(dst, big[1 : 1 + src.numel()])  # the plane kernel takes any view
    torch.cuda.synchronize()


def test_vmult_accumulates_like_reference(mfhn):
    """cell_loop passes no zero flag (benchmark_03.h:352): dst += A src."""
    import torch

    dh, mf, lay = _case(mfhn, "quadrant", 3, "serial", 2)
    x = _src(lay, "random")
    op = mfhn.LaplaceOperator(mf)
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    src.copy_(torch.from_numpy(x))
    op.vmult(dst, src)
    op.vmult(dst, src)
    once = operators.vmult_fast(lay, x)
    assert np.abs(dst.cpu().numpy() - 2 * once).max() / np.abs(once).max() < 1e-12
    op.vmult(dst, src, zero_dst=True)
    assert np.abs(dst.cpu().numpy() - once).max() / np.abs(once).max() < 1e-12


def test_constant_is_in_null_space_large(mfhn):
    """A 1 = 0 (src == 1 is what benchmark_01.h:510-511 times) at a size the oracle
    would not finish quickly: size-independent property."""
    import torch

    tria = mfhn.Triangulation("annulus", 7, "p4est")
    dh = mfhn.DoFHandler(tria, 4)
    mf = mfhn.MatrixFree(dh)
    for kernel in KERNELS:
        op = mfhn.LaplaceOperator(mf, kernel=kernel)
        src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
        src.fill_(1.0)
        op.vmult(dst, src)
        assert dst.abs().max().item() < 1e-11


def test_symmetry_and_energy_large(mfhn):
    """x^T A y = y^T A x and u^T A u = int |grad p|^2 for a polynomial p of degree <= k
    on a mesh with hanging nodes (exact for the conforming space)."""
    import torch

    tria = mfhn.Triangulation("quadrant", 5, "p4est")
    k = 3
    dh = mfhn.DoFHandler(tria, k)
    mf = mfhn.MatrixFree(dh)
    pts = dh.support_points()
    X, Y, Z = pts[:, 0], pts[:, 1], pts[:, 2]
    p = X * X + Y * Z + 0.5 * X * Y * Z + X * Y * Y
    # int over (-1,1)^3 of |grad p|^2 with grad p = (2x + yz/2 + y^2, z + xz/2 + 2xy, y + xy/2)
    import sympy as sp

    x, y, z = sp.symbols("x y z")
    pe = x * x + y * z + sp.Rational(1, 2) * x * y * z + x * y * y
    exact = float(sp.integrate(sum(sp.diff(pe, v) ** 2 for v in (x, y, z)), (x, -1, 1), (y, -1, 1), (z, -1, 1)))
    rng = np.random.default_rng(12345)
    a, b = rng.uniform(-1, 1, dh.n_dofs()), rng.uniform(-1, 1, dh.n_dofs())
    for kernel in KERNELS:
        op = mfhn.LaplaceOperator(mf, kernel=kernel)
        src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()

        def A(v):
            src.copy_(torch.from_numpy(v))
            op.vmult(dst, src, zero_dst=True)
            return dst.cpu().numpy()

        Ap = A(p)
        # hanging entries of dst are never written and hanging entries of src never read:
        # the energy uses the live entries only
        assert abs(p @ Ap - exact) / exact < 1e-12, (kernel, p @ Ap, exact)
        Aa, Ab = A(a), A(b)
        live = Aa != 0
        assert abs(b[live] @ Aa[live] - a[live] @ Ab[live]) / abs(b[live] @ Aa[live]) < 1e-10


def _metric(J):
    """det J J^-1 J^-T of [..., 3, 3] Jacobians as the six components (xx,xy,xz,yy,yz,zz)."""
    Ji = np.linalg.inv(J)
    M = np.linalg.det(J)[..., None, None] * (Ji @ np.swapaxes(Ji, -1, -2))
    return np.stack([M[..., 0, 0], M[..., 0, 1], M[..., 0, 2], M[..., 1, 1], M[..., 1, 2], M[..., 2, 2]], axis=-1)


def _oracle_layout_in_operator_order(mfhn, mf, lay, geo, L, flavour):
    """The oracle's layout with the cells in the operator's (reordered) cell order."""
    import copy

    perm = {tuple(c): i for i, c in enumerate(lay.cells.tolist())}
    cells = mfhn.Triangulation(geo, L, flavour).cells()[mf.cell_ids]
    order = np.array([perm[tuple(c)] for c in cells.tolist()])
    lo = copy.copy(lay)
    lo.cells, lo.dof_indices, lo.kinds, lo.masks, lo.h = lay.cells[order], lay.dof_indices[order], lay.kinds[order], lay.masks[order], lay.h[order]
    assert np.array_equal(lo.dof_indices, mf.dof_indices) and np.array_equal(lo.masks, mf.masks)
    return lo


@pytest.mark.parametrize("k", [2, 4, 5])
@pytest.mark.parametrize("number", ["double", "float"])
def test_affine_geometry_sheared_and_rotated(mfhn, k, number):
    """MFHN_GEOM_AFFINE with full 3x3 Jacobians (shear + rotation + anisotropic scaling, different per cell): all six
    components of det J J^-1 J^-T are non-zero.  Against the numpy restatement of evaluate / submit_gradient / integrate
    (benchmark_03.h:284-290) with the same coefficient at every quadrature point."""
    import torch

    from oracle import fe1d

    geo, L = ("annulus", 5) if k <= 4 else ("quadrant", 4)
    dh, mf, lay = _case(mfhn, geo, L, "serial", k)
    lo = _oracle_layout_in_operator_order(mfhn, mf, lay, geo, L, "serial")
    n = k + 1
    rng = np.random.default_rng(3)
    th = rng.uniform(0, 2 * np.pi, mf.n_cells)
    Rz = np.zeros((mf.n_cells, 3, 3))
    Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1], Rz[:, 2, 2] = np.cos(th), -np.sin(th), np.sin(th), np.cos(th), 1.0
    shear = np.array([[1.0, 0.35, -0.2], [0.1, 1.2, 0.25], [-0.15, 0.3, 0.8]])
    J = mf.h[:, None, None] * (Rz @ (shear[None] + rng.uniform(-0.05, 0.05, (mf.n_cells, 3, 3))))
    assert (np.linalg.det(J) > 0).all()
    w = fe1d.shape_data(k).qw
    w3 = (w[:, None, None] * w[None, :, None] * w[None, None, :]).ravel()
    G = np.einsum("cm,q->cmq", _metric(J), w3)
    assert min(np.abs(G[:, comp]).max() for comp in (1, 2, 4)) > 1e-3  # the off-diagonal terms are live
    x = _src(lay, "random")
    ref = operators.vmult_general(lo, x, G)
    for kernel in ("auto", "qpoint"):
        op = mfhn.LaplaceOperator(mf, number=number, kernel=kernel, geometry=J)
        src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
        src.copy_(torch.from_numpy(x).to(src.dtype))
        op.vmult(dst, src)
        err = np.abs(dst.cpu().numpy().astype(np.float64) - ref).max() / np.abs(ref).max()
        assert err < TOL[number], (k, number, kernel, err)


@pytest.mark.parametrize("k", [2, 3, 4, 6])
def test_general_geometry_full_tensor(mfhn, k):
    """MFHN_GEOM_GENERAL with a deformation whose Jacobian is a full matrix at every quadrature point
    (x = X + eps d(X), d mixes the coordinates): xy / xz / yz coefficients are non-zero and vary inside the cell --
    the data class of TestHighOrderMapping (benchmark_01.h:225-242, benchmark_03.h:284-290)."""
    import torch

    from oracle import fe1d

    geo, L = ("annulus", 5) if k <= 4 else ("quadrant", 4)
    dh, mf, lay = _case(mfhn, geo, L, "serial", k)
    lo = _oracle_layout_in_operator_order(mfhn, mf, lay, geo, L, "serial")
    n = k + 1
    sd = fe1d.shape_data(k)
    q, w = sd.qpts, sd.qw
    w3 = (w[:, None, None] * w[None, :, None] * w[None, None, :]).ravel()
    hh = lo.h
    org = -1.0 + lo.cells[:, 1:4] * hh[:, None]
    qx, qy, qz = np.tile(q, n * n), np.tile(np.repeat(q, n), n), np.repeat(q, n * n)
    X, Y, Z = (org[:, d, None] + hh[:, None] * qq[None, :] for d, qq in enumerate((qx, qy, qz)))
    eps, pi = 2e-2, np.pi
    # d = (sin(pi Y) + sin(pi Z), sin(pi X) cos(pi Z), sin(pi (X + Y))): gradient of the deformation at the quadrature points
    D = np.zeros(X.shape + (3, 3))
    D[..., 0, 1], D[..., 0, 2] = pi * np.cos(pi * Y), pi * np.cos(pi * Z)
    D[..., 1, 0], D[..., 1, 2] = pi * np.cos(pi * X) * np.cos(pi * Z), -pi * np.sin(pi * X) * np.sin(pi * Z)
    D[..., 2, 0] = D[..., 2, 1] = pi * np.cos(pi * (X + Y))
    J = hh[:, None, None, None] * (np.eye(3) + eps * D)
    assert (np.linalg.det(J) > 0).all()
    G = np.moveaxis(_metric(J), -1, 1) * w3[None, None, :]  # [cell][6][q]
    assert min(np.abs(G[:, comp]).max() for comp in (1, 2, 4)) > 1e-4
    x = _src(lay, "random")
    ref = operators.vmult_general(lo, x, G)
    for kernel in ("auto", "qpoint"):
        op = mfhn.LaplaceOperator(mf, kernel=kernel, geometry=G)
        src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
        src.copy_(torch.from_numpy(x))
        op.vmult(dst, src)
        y = dst.cpu().numpy()
        assert np.abs(y - ref).max() / np.abs(ref).max() < 1e-12, (k, kernel)
    # symmetry and null space survive the full tensor
    b = np.random.default_rng(9).uniform(-1, 1, lay.n_dofs)
    Ab = operators.vmult_general(lo, b, G)
    live = ref != 0
    assert abs(b[live] @ y[live] - x[live] @ Ab[live]) < 1e-10 * abs(b[live] @ y[live])
    src.fill_(1.0)
    op.vmult(dst, src, zero_dst=True)
    assert dst.abs().max().item() < 1e-10


def test_affine_geometry_matches_cartesian(mfhn):
    """Affine path with J = h I must reproduce the Cartesian result."""
    import torch

    dh, mf, lay = _case(mfhn, "annulus", 5, "serial", 2)
    x = _src(lay, "random")
    ref = operators.vmult_fast(lay, x)
    J = np.einsum("c,ij->cij", mf.h, np.eye(3))
    op = mfhn.LaplaceOperator(mf, kernel="qpoint", geometry=J)
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    src.copy_(torch.from_numpy(x))
    op.vmult(dst, src)
    assert np.abs(dst.cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-12


def test_hanging_node_kernel_alone(mfhn):
    """FEEvaluationHangingNodesFactory::apply analogue (benchmark_00_likwid.cc:56-59)."""
    import torch

    for k in (1, 2, 4, 7):
        dh, mf, lay = _case(mfhn, "annulus", 5 if k < 7 else 4, "serial", k)
        op = mfhn.LaplaceOperator(mf)
        n = k + 1
        rng = np.random.default_rng(7)
        vals = rng.uniform(-1, 1, (mf.n_cells, n, n, n))
        # mf cells are reordered: recover kinds from the masks
        kinds = np.array([dofs.decompress(int(m)) for m in mf.masks], dtype=np.uint16)
        for transpose in (False, True):
            d = torch.from_numpy(vals.copy()).cuda()
            op.apply_hanging_node_constraints(d, transpose)
            ref = operators.hn_apply(vals.copy(), kinds, k, transpose)
            assert np.abs(d.cpu().numpy() - ref).max() < 1e-13


def test_error_behaviour(mfhn):
    """Unsupported inputs raise like the reference's AssertThrow paths."""
    with pytest.raises(mfhn.MfhnError):
        mfhn.Triangulation("torus", 3)  # "Unknown geometry type!" benchmark_03.h:404
    tria = mfhn.Triangulation("quadrant", 2, "serial")
    with pytest.raises(mfhn.MfhnError):
        mfhn.DoFHandler(tria, 9)  # degree dispatch default case, benchmark_01.cc:64-66
    dh = mfhn.DoFHandler(tria, 6)
    mf = mfhn.MatrixFree(dh)
    with pytest.raises(mfhn.MfhnError):
        mfhn.LaplaceOperator(mf, kernel="bulk")  # not available for this degree
    op = mfhn.LaplaceOperator(mf)
    v = op.initialize_dof_vector()
    with pytest.raises(mfhn.MfhnError):
        op.vmult(v, v)


def test_cpp_driver_over_c_abi(mfhn):
    """examples/benchmark_03 (C++ host code over the C ABI, laid out like the reference's
    benchmark_03 run()) reproduces the oracle's |A src|_2 for src = sum sin x_d."""
    import os
    import re
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "benchmark_03")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(root, "examples")])
    out = subprocess.run([exe, "annulus", "3", "5", "5"], check=True, capture_output=True, text=True).stdout
    got = float(re.search(r"n_repetitions = ([0-9.e+-]+)", out).group(1))
    lay = dofs.setup(mesh.create("annulus", 5, "p4est"), 3)
    ref = np.linalg.norm(operators.vmult_fast(lay, np.sin(lay.support_points).sum(axis=1)))
    assert abs(got - ref) / ref < 1e-10
    cols = out.splitlines()[1].split()
    assert int(cols[0]) == 1 and int(cols[4]) == lay.n_cells and int(cols[6]) == lay.n_dofs and int(cols[5]) == int((lay.masks != 0).sum())
    # the reference's `mpirun -np 2`: forked ranks, one GPU each, NCCL ghost exchange (needs two devices)
    import torch

    if torch.cuda.device_count() >= 2:
        out = subprocess.run([exe, "annulus", "3", "5", "5", "2"], check=True, capture_output=True, text=True).stdout
        got = float(re.search(r"n_repetitions = ([0-9.e+-]+)", out).group(1))
        assert abs(got - ref) / ref < 1e-10


def test_host_vector_entry_point(mfhn):
    """mfhn_op_vmult_host[_slot]: the call a host-vector caller binds (LaplaceOperator<...,Host>::vmult,
    benchmark_03.h:237-241): same result as the device-vector path, both staging slots, accumulate and zero."""
    import torch

    dh, mf, lay = _case(mfhn, "annulus", 5, "p4est", 2)
    x = _src(lay, "random")
    ref = operators.vmult_fast(lay, x)
    op = mfhn.LaplaceOperator(mf)
    hs = torch.from_numpy(x.copy()).pin_memory()
    for slot in (0, 1):
        hd = torch.zeros(lay.n_dofs, dtype=torch.float64).pin_memory()
        op.vmult_host(hd, hs, zero_dst=True, slot=slot)
        torch.cuda.synchronize()
        assert np.abs(hd.numpy() - ref).max() / np.abs(ref).max() < 1e-12
        op.vmult_host(hd, hs, zero_dst=False, slot=slot)  # accumulates like the reference
        torch.cuda.synchronize()
        assert np.abs(hd.numpy() - 2 * ref).max() / np.abs(ref).max() < 1e-12
    with pytest.raises(mfhn.MfhnError):
        op.vmult_host(hd, hs, slot=2)


@pytest.mark.parametrize("geo,L,k", [("annulus", 5, 2), ("quadrant", 3, 4), ("quadrant", 2, 7)])
def test_diagonal_and_jacobi_cg(mfhn, geo, L, k):
    """Extension (BASELINE.json config 5; the reference has no solver): the operator diagonal equals
    diag(C^T K C) of the general-purpose oracle, and point-Jacobi CG solves a consistent system."""
    import torch

    t = mesh.create(geo, L, "serial")
    lay = dofs.setup(t, k)
    O1 = operators.GeneralOperator(t, lay)
    A = (O1.C.T @ O1.K @ O1.C).tocsr()
    tria = mfhn.Triangulation(geo, L, "serial")
    dh = mfhn.DoFHandler(tria, k)
    mf = mfhn.MatrixFree(dh)
    op = mfhn.LaplaceOperator(mf)
    diag = op.compute_diagonal()
    ref = A.diagonal()
    assert np.abs(diag.cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-12
    assert (diag.cpu().numpy()[O1.is_hanging] == 0).all()
    # consistent right-hand side b = A x*, x* random on the live DoFs
    live = ~O1.is_hanging
    xs = np.zeros(lay.n_dofs)
    xs[live] = np.random.default_rng(5).uniform(-1, 1, live.sum())
    b = op.initialize_dof_vector()
    b.copy_(torch.from_numpy(A @ xs))
    x = op.initialize_dof_vector()
    its_jacobi, hist = mfhn.solve_cg(op, x, b, diag=diag, rel_tol=1e-10, max_iter=2000)
    # the library's own inverse diagonal (mfhn_op_inverse_diagonal) and the device-time split
    inv = mfhn.inverse_diagonal(op)
    assert np.allclose(inv.cpu().numpy()[live], 1.0 / ref[live], rtol=1e-12) and (inv.cpu().numpy()[O1.is_hanging] == 0).all()
    x3 = op.initialize_dof_vector()
    its3, hist3, split = mfhn.solve_cg(op, x3, b, inverse=inv, rel_tol=1e-10, max_iter=2000, check_every=10, timings=True)
    assert its_jacobi <= its3 <= its_jacobi + 10 and split["ms_vmult"] > 0 and split["ms_vector_ops"] > 0
    assert hist[-1] <= 1e-10 * hist[0] and its_jacobi < 2000
    xn = x.cpu().numpy()
    assert np.abs(A @ xn - A @ xs).max() <= 1e-8 * np.abs(A @ xs).max()
    d = (xn - xs)[live]
    assert np.abs(d - d.mean()).max() < 1e-6 * np.abs(xs).max()  # equal up to the constant null space
    x2 = op.initialize_dof_vector()
    its_plain, _ = mfhn.solve_cg(op, x2, b, diag=None, rel_tol=1e-10, max_iter=4000)
    assert its_jacobi <= its_plain  # the preconditioner pays off on the graded mesh


def test_general_per_quadrature_point_geometry(mfhn):
    """MFHN_GEOM_GENERAL: one symmetric coefficient JxW J^-1 J^-T per quadrature point -- the data class of the
    reference's TestHighOrderMapping (benchmark_01.h:225-242).  (i) Cartesian coefficients reproduce the Cartesian
    operator; (ii) the smooth deformation x = X + 1e-2 sin(pi X) matches the numpy restatement; (iii) the operator
    stays symmetric and keeps the constants in its null space."""
    import torch

    from oracle import fe1d

    k = 3
    n = k + 1
    dh, mf, lay = _case(mfhn, "annulus", 5, "serial", k)
    # the operator's cells are reordered: build the oracle layout in the same order
    import copy

    perm = {tuple(c): i for i, c in enumerate(lay.cells.tolist())}
    cells = mfhn.Triangulation("annulus", 5, "serial").cells()[mf.cell_ids]
    order = np.array([perm[tuple(c)] for c in cells.tolist()])
    lo = copy.copy(lay)
    lo.cells, lo.dof_indices, lo.kinds, lo.masks, lo.h = lay.cells[order], lay.dof_indices[order], lay.kinds[order], lay.masks[order], lay.h[order]
    assert np.array_equal(lo.dof_indices, mf.dof_indices) and np.array_equal(lo.masks, mf.masks)
    sd = fe1d.shape_data(k)
    q, w = sd.qpts, sd.qw
    w3 = (w[:, None, None] * w[None, :, None] * w[None, None, :]).ravel()  # [z][y][x] -> lexicographic x fastest
    x = _src(lay, "random")

    def run(G):
        op = mfhn.LaplaceOperator(mf, kernel="qpoint", geometry=G)
        src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
        src.copy_(torch.from_numpy(x))
        op.vmult(dst, src)
        return dst.cpu().numpy(), op

    # (i) Cartesian: JxW J^-1 J^-T = w_q h I
    G = np.zeros((mf.n_cells, 6, n ** 3))
    for comp in (0, 3, 5):
        G[:, comp, :] = lo.h[:, None] * w3[None, :]
    ref = operators.vmult_fast(lo, x)
    y, _ = run(G)
    assert np.abs(y - ref).max() / np.abs(ref).max() < 1e-12
    assert np.abs(operators.vmult_general(lo, x, G) - ref).max() / np.abs(ref).max() < 1e-12
    # (ii) x = X + eps sin(pi X) componentwise: J = h diag(1 + eps pi cos(pi X_d)) at the quadrature points
    eps = 1e-2
    hh = lo.h
    org = -1.0 + lo.cells[:, 1:4] * hh[:, None]
    qx = np.tile(q, n * n)
    qy = np.tile(np.repeat(q, n), n)
    qz = np.repeat(q, n * n)
    X = [org[:, d, None] + hh[:, None] * qq[None, :] for d, qq in enumerate((qx, qy, qz))]
    Jd = [hh[:, None] * (1 + eps * np.pi * np.cos(np.pi * X[d])) for d in range(3)]
    det = Jd[0] * Jd[1] * Jd[2]
    G = np.zeros((mf.n_cells, 6, n ** 3))
    for comp, d in ((0, 0), (3, 1), (5, 2)):
        G[:, comp, :] = w3[None, :] * det / (Jd[d] * Jd[d])
    ref = operators.vmult_general(lo, x, G)
    y, op = run(G)
    assert np.abs(y - ref).max() / np.abs(ref).max() < 1e-12
    # (iii) symmetry and null space
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    src.fill_(1.0)
    op.vmult(dst, src)
    assert dst.abs().max().item() < 1e-11
    b = np.random.default_rng(9).uniform(-1, 1, lay.n_dofs)
    Ab = operators.vmult_general(lo, b, G)
    live = ref != 0
    assert abs(b[live] @ y[live] - x[live] @ Ab[live]) < 1e-10 * abs(b[live] @ y[live])


@pytest.mark.parametrize("k", [2, 4, 6])
def test_hn_mask_strategy_and_dg_copy(mfhn, k):
    """(i) MFHN_HN_MASK (every warp takes the interpolation passes, per-lane predicates: the reference's "mask"
    vectorisation type, benchmark_01.cc:70-116) gives the same vmult as the default branch strategy, for every fast
    kernel.  (ii) DG (C) stage (benchmark_01.cc:189-199): dst += W^T W src on cell-local values against the oracle's
    interpolation matrices."""
    import torch

    geo, L = ("annulus", 5) if k <= 4 else ("quadrant", 4)
    dh, mf, lay = _case(mfhn, geo, L, "serial", k)
    x = _src(lay, "random")
    ref = operators.vmult_fast(lay, x)
    op = mfhn.LaplaceOperator(mf)
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    src.copy_(torch.from_numpy(x))
    for kern in ("plane", "bulk", "runs"):
        try:
            op.set_kernel(kern)
            op.set_hn_strategy("mask")
            op.vmult(dst, src, zero_dst=True)
        except mfhn.MfhnError:
            continue
        finally:
            op.set_hn_strategy("branch")
        y = dst.cpu().numpy()
        assert np.abs(y - ref).max() / np.abs(ref).max() < 1e-12, (k, kern)
    # general-purpose constraint algorithm (use_fast_hanging_node_algorithm = false, benchmark_01.h:286-293): weighted
    # constraint rows in the gather / scatter instead of the interpolation passes -- same operator
    for number in ("double", "float"):
        opr = mfhn.LaplaceOperator(mf, number=number, kernel="qpoint_rows")
        s2, d2 = opr.initialize_dof_vector(), opr.initialize_dof_vector()
        s2.copy_(torch.from_numpy(x).to(s2.dtype))
        opr.vmult(d2, s2)
        assert np.abs(d2.double().cpu().numpy() - ref).max() / np.abs(ref).max() < TOL[number], (k, number)
        assert opr.query("constraint_row_entries") > (k + 1) ** 3
        opr.set_apply_constraints(False)
        opr.vmult(d2, s2, zero_dst=True)
        ref_nc = operators.vmult_fast(lay, x, apply_constraints=False)
        assert np.abs(d2.double().cpu().numpy() - ref_nc).max() / np.abs(ref_nc).max() < TOL[number]
    # DG (C): cell-local values, constraints on / off
    n, n3 = k + 1, (k + 1) ** 3
    rng = np.random.default_rng(3)
    v = rng.uniform(-1, 1, (mf.n_cells, n3))
    kinds = np.array([dofs.decompress(int(m)) for m in mf.masks], dtype=np.uint16)
    want = operators.hn_apply(operators.hn_apply(v.reshape(-1, n, n, n).copy(), kinds, k, False), kinds, k, True).reshape(-1, n3)
    sv, dv = torch.from_numpy(v).cuda().reshape(-1), torch.zeros(mf.n_cells * n3, dtype=torch.float64, device="cuda")
    op.dg_copy(dv, sv)
    assert np.abs(dv.cpu().numpy().reshape(-1, n3) - want).max() < 1e-12
    op.set_apply_constraints(False)
    op.dg_copy(dv, sv)
    assert np.abs(dv.cpu().numpy().reshape(-1, n3) - want - v).max() < 1e-12
