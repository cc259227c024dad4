"""Conjugate gradients with a point-Jacobi preconditioner on top of the
operator (BASELINE.json config 5).  EXTENSION: the reference contains no solver
(SURVEY.md 0.5); this is the step either side of vmult in a real solve --
vmult + vector updates + dot-product all-reduce -- kept deliberately small.
Vector updates and dot products are plain torch device ops, the all-reduce is
torch.distributed; the operator and its diagonal are the CUDA kernels of this
package."""
from __future__ import annotations


def solve_cg(op, x, b, diag=None, rel_tol=1e-8, max_iter=1000, group=None):
    """Solves A x = b for the (singular, positive semi-definite) Laplace operator
    `op` starting from x; b must be consistent (orthogonal to the constants).
    Only the locally owned entries of the vectors enter the dot products.
    Returns (iterations, list of residual norms)."""
    import torch

    n_owned = op.n_owned
    dist = None
    if op._comm is not None:
        import torch.distributed as dist_mod

        dist = dist_mod

    def dot(u, v):
        s = torch.dot(u[:n_owned], v[:n_owned]).reshape(1)
        if dist is not None:
            dist.all_reduce(s, group=group)
        return s

    inv = None
    if diag is not None:
        inv = torch.where(diag != 0, 1.0 / torch.where(diag != 0, diag, torch.ones_like(diag)), torch.zeros_like(diag))
    r = op.initialize_dof_vector()
    op.vmult(r, x, zero_dst=True)
    r[:n_owned] = b[:n_owned] - r[:n_owned]
    r[n_owned:] = 0
    z = r * inv if inv is not None else r.clone()
    p = z.clone()
    Ap = op.initialize_dof_vector()
    rz = dot(r, z)
    r0 = float(torch.sqrt(dot(r, r)))
    history = [r0]
    if r0 == 0.0:
        return 0, history
    for it in range(1, max_iter + 1):
        p[n_owned:] = 0
        op.vmult(Ap, p, zero_dst=True)
        alpha = rz / dot(p, Ap)
        x[:n_owned] += alpha * p[:n_owned]
        r[:n_owned] -= alpha * Ap[:n_owned]
        res = float(torch.sqrt(dot(r, r)))
        history.append(res)
        if res <= rel_tol * r0:
            return it, history
        z = r * inv if inv is not None else r
        rz_new = dot(r, z)
        p[:n_owned] = z[:n_owned] + (rz_new / rz) * p[:n_owned]
        rz = rz_new
    return max_iter, history
