"""Krylov solvers over block-sparse vectors (oracle; test infrastructure).

Restates the two KrylovKit 0.9.5 entry points the hot path uses (Manifest.toml:548; not
vendored):
  * `eigsolve(f, x0, 1, :SR, Lanczos(krylovdim=30, tol, maxiter))` -> `lanczos_lowest`
  * `linsolve(f, b, GMRES(krylovdim=30, tol, maxiter))`              -> `gmres`
Vectors are any container with `.blocks` (dict of ndarrays) and `.weight(key)`; the inner
product is TensorKit's  <x,y> = sum_blocks weight * tr(x^T y)  (SURVEY.md App. A).
"""
from __future__ import annotations

import numpy as np


def vdot(x, y) -> float:
    return float(sum(x.weight(k) * np.vdot(x.blocks[k], y.blocks[k]) for k in x.blocks))


def vnorm(x) -> float:
    return float(np.sqrt(max(vdot(x, x), 0.0)))


def vaxpy(alpha, x, y):
    for k in x.blocks:
        y.blocks[k] += alpha * x.blocks[k]
    return y


def vscale(x, alpha):
    for k in x.blocks:
        x.blocks[k] *= alpha
    return x


def vcopy(x):
    return x.copy()


def lanczos_lowest(apply, x0, tol=1e-10, krylovdim=30, maxiter=100):
    """Lowest eigenpair of the symmetric operator `apply` (Lanczos with full
    re-orthogonalisation, explicit restart from the Ritz vector).  Returns (eval, evec, info)
    with info = dict(converged, residual, applies)."""
    x = vcopy(x0)
    nrm = vnorm(x)
    if nrm == 0.0:
        raise ValueError("lanczos: zero start vector")
    vscale(x, 1.0 / nrm)
    applies = 0
    theta, res = 0.0, np.inf
    for _ in range(maxiter):
        V = [x]
        alphas, betas = [], []
        w = apply(V[0])
        applies += 1
        for j in range(krylovdim):
            a = vdot(V[j], w)
            alphas.append(a)
            vaxpy(-a, V[j], w)
            if j > 0:
                vaxpy(-betas[j - 1], V[j - 1], w)
            for _pass in range(2):                      # full re-orthogonalisation (CGS2-like)
                for q in V:
                    vaxpy(-vdot(q, w), q, w)
            b = vnorm(w)
            Tm = np.diag(alphas) + np.diag(betas, 1) + np.diag(betas, -1)
            ev, evec = np.linalg.eigh(Tm)
            theta, y = ev[0], evec[:, 0]
            res = abs(b * y[-1])
            if res < tol or b < 1e-14 or j == krylovdim - 1:
                break
            betas.append(b)
            V.append(vscale(w, 1.0 / b))
            w = apply(V[-1])
            applies += 1
        xnew = vscale(vcopy(V[0]), y[0])
        for q, c in zip(V[1:], y[1:]):
            vaxpy(c, q, xnew)
        vscale(xnew, 1.0 / vnorm(xnew))
        x = xnew
        if res < tol:
            return theta, x, dict(converged=True, residual=res, applies=applies)
    return theta, x, dict(converged=False, residual=res, applies=applies)


def gmres(apply, b, x0=None, tol=1e-10, krylovdim=30, maxiter=100):
    """Restarted GMRES for apply(x) = b.  Returns (x, info)."""
    x = vscale(vcopy(b), 0.0) if x0 is None else vcopy(x0)
    bnorm = vnorm(b)
    if bnorm == 0.0:
        return x, dict(converged=True, residual=0.0, applies=0)
    applies = 0
    res = np.inf
    for _ in range(maxiter):
        r = vaxpy(-1.0, apply(x), vcopy(b))
        applies += 1
        beta = vnorm(r)
        res = beta / bnorm
        if res < tol:
            return x, dict(converged=True, residual=res, applies=applies)
        V = [vscale(r, 1.0 / beta)]
        H = np.zeros((krylovdim + 1, krylovdim))
        k_used = 0
        for j in range(krylovdim):
            w = apply(V[j])
            applies += 1
            for i, q in enumerate(V):
                H[i, j] = vdot(q, w)
                vaxpy(-H[i, j], q, w)
            for q in V:                                  # second pass (CGS2)
                c = vdot(q, w)
                H[V.index(q), j] += c
                vaxpy(-c, q, w)
            H[j + 1, j] = vnorm(w)
            k_used = j + 1
            e1 = np.zeros(j + 2)
            e1[0] = beta
            y, *_ = np.linalg.lstsq(H[:j + 2, :j + 1], e1, rcond=None)
            res = np.linalg.norm(H[:j + 2, :j + 1] @ y - e1) / bnorm
            if res < tol or H[j + 1, j] < 1e-14:
                break
            V.append(vscale(w, 1.0 / H[j + 1, j]))
        for q, c in zip(V[:k_used], y):
            vaxpy(c, q, x)
        if res < tol:
            return x, dict(converged=True, residual=res, applies=applies)
    return x, dict(converged=False, residual=res, applies=applies)
