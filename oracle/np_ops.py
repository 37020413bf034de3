"""Literal numpy/scipy transcription of the reference's op bodies -- TEST INFRASTRUCTURE ONLY.

Follows lib/models.py line by line (same transposes, reshapes, stacking), with scipy's fp32 CSR SpMM
standing in for tf.sparse_tensor_dense_matmul and numpy matmul for tf.matmul.  Used to pin
oracle/cape_oracle.py, and -- being the closest thing to the reference's "TF1 CPU path" arithmetic that
can run here -- as a per-op CPU baseline.  Also holds an independent float64 dense-polynomial
formulation of the Chebyshev conv used as a known-answer generator.
"""
import numpy as np
import scipy.sparse as sp

from .cape_oracle import rescale_L


def chebyshev5_np(x, L, W, K):
    """lib/models.py:69-103.  x [N,M,Fin] fp32, L scipy sparse (un-rescaled Laplacian), W [Fin*K, Fout]."""
    N, M, Fin = x.shape
    L = sp.csr_matrix(L)                      # :74
    L = rescale_L(L, lmax=2)                  # :75
    x0 = np.transpose(x, (1, 2, 0))           # :81  M x Fin x N
    x0 = np.reshape(x0, (M, Fin * N))         # :82
    xs = x0[None]                             # :83

    def concat(xs, x_):
        return np.concatenate([xs, x_[None]], axis=0)   # :85-87

    if K > 1:
        x1 = L.dot(x0)                        # :91
        xs = concat(xs, x1)
    for _ in range(2, K):
        x2 = 2 * L.dot(x1) - x0               # :94
        xs = concat(xs, x2)
        x0, x1 = x1, x2
    xs = np.reshape(xs, (K, M, Fin, N))       # :97
    xs = np.transpose(xs, (3, 1, 2, 0))       # :98
    xs = np.reshape(xs, (N * M, Fin * K))     # :99
    y = xs @ W                                # :102
    return np.reshape(y, (N, M, -1))          # :103


def b1leakyrelu_np(x, b, alpha=0.2):
    """lib/models.py:105-109."""
    y = x + np.reshape(b, (1, 1, -1))
    return np.where(y > 0, y, alpha * y).astype(x.dtype)


def poolwT_np(x, S):
    """lib/models.py:129-152."""
    Mp = S.shape[0]
    N, M, Fin = x.shape
    S = sp.csr_matrix(S)
    xt = np.transpose(x, (1, 2, 0)).reshape(M, Fin * N)
    xt = S.dot(xt)
    xt = np.reshape(xt, (Mp, Fin, N))
    return np.transpose(xt, (2, 0, 1))


def chebyshev_dense_f64(x, L, W, K):
    """Independent formulation: y = sum_k T_k(L~) x W[k::K] with dense float64 polynomials."""
    Lt = rescale_L(sp.csr_matrix(L), lmax=2).astype(np.float64).toarray()
    M = Lt.shape[0]
    T = [np.eye(M), Lt]
    for _ in range(2, K):
        T.append(2 * Lt @ T[-1] - T[-2])
    x = x.astype(np.float64)
    W = W.astype(np.float64)
    y = 0
    for k in range(K):
        y = y + np.einsum("uv,nvf,fo->nuo", T[k], x, W[k::K])
    return y
