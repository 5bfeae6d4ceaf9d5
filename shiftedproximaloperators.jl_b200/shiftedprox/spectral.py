"""Thresholding stage of the reference's spectral operators, given an SVD.

ShiftedRank / ShiftedNuclearnorm / ShiftedCappedl1 `prox!` (src/shiftedRank.jl:68-84, src/shiftedNuclearnorm.jl:68-81,
src/shiftedCappedl1.jl:68-86) are SVD-bound and out of scope as operators; what is on this side of the boundary is
the elementwise work around the factorisation: `sol = q + xk + sj`, the thresholding of the singular values with the
column scaling of U, and `y = A - (xk + sj)`.  The SVD (LAPACK `gesdd` in the reference, src/psvd.jl) and the `mul!`
stay library calls: `prox_spectral_` below uses torch.linalg.svd (cuSOLVER) and torch.matmul (cuBLAS) for them.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from . import _p, _SUF, _vec, context

KINDS = {"rank": 0, "nuclear": 1, "cappedl1": 2}


def spectral_sol_(a_out: torch.Tensor, xk: torch.Tensor, sj: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """a_out = (q + xk) + sj   (`ψ.sol .= q .+ ψ.xk .+ ψ.sj`, shiftedRank.jl:69)."""
    _vec(q, name="q"); _vec(xk, q, "xk"); _vec(sj, q, "sj"); _vec(a_out, q, "a_out")
    L.call(f"spx_spectral_sol_{_SUF[q.dtype]}", context(q.device), C.c_int64(q.numel()), _p(a_out), _p(xk), _p(sj), _p(q))
    return a_out


def spectral_threshold_(U: torch.Tensor, S: torch.Tensor, kind: str, lam: float, sigma: float, theta: float = 0.0):
    """In place: S' = threshold(S) and U[:, i] *= S'_i.  U is an m x k COLUMN-MAJOR matrix (a torch tensor whose
    transpose is contiguous, e.g. `torch.linalg.svd(A)[0].T.contiguous().T`), S its k singular values."""
    if U.dim() != 2 or U.stride(0) != 1:
        raise ValueError("U must be column-major (stride(0) == 1)")
    m, k = U.shape
    if S.numel() != k or S.dtype != U.dtype:
        raise ValueError("S must hold one singular value per column of U, in U's element type")
    L.call(f"spx_spectral_threshold_{_SUF[U.dtype]}", context(U.device), C.c_int32(KINDS[kind]), C.c_int64(m), C.c_int64(k),
           C.c_void_p(U.data_ptr()), C.c_int64(U.stride(1) if k > 1 else max(m, 1)), _p(S), C.c_double(lam),
           C.c_double(sigma), C.c_double(theta))
    return U, S


def spectral_finish_(y: torch.Tensor, a: torch.Tensor, xk: torch.Tensor, sj: torch.Tensor) -> torch.Tensor:
    """y = a - (xk + sj)   (shiftedRank.jl:82)."""
    _vec(a, name="a"); _vec(xk, a, "xk"); _vec(sj, a, "sj"); _vec(y, a, "y")
    L.call(f"spx_spectral_finish_{_SUF[a.dtype]}", context(a.device), C.c_int64(a.numel()), _p(y), _p(a), _p(xk), _p(sj))
    return y


def prox_spectral_(y, kind, shape, xk, sj, q, lam, sigma, theta=0.0):
    """prox! of ShiftedRank ("rank") / ShiftedNuclearnorm ("nuclear") / ShiftedCappedl1 ("cappedl1") on vectors that are
    the column-major image of a `shape` = (m, n) matrix.  The factorisation and the product are LIBRARY calls
    (cuSOLVER / cuBLAS through torch) and not part of the path; the three stages around them are libshiftedprox's."""
    m, n = shape
    sol = torch.empty_like(q)
    spectral_sol_(sol, xk, sj, q)
    A = sol.view(n, m).T  # column-major m x n
    U, S, Vt = torch.linalg.svd(A, full_matrices=False)
    U = U.T.contiguous().T  # column-major
    S = S.contiguous()
    spectral_threshold_(U, S, kind, lam, sigma, theta)
    out = torch.matmul(U, Vt)  # `mul!(ψ.h.A, ψ.h.F.U, ψ.h.F.Vt)`
    a = out.T.contiguous().view(-1)  # back to the column-major vector
    return spectral_finish_(y, a, xk, sj)
