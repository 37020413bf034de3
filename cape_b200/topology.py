"""Fixed SMPL mesh hierarchy: loading, Laplacians, operator composition, ELL packing.

Host-side counterpart of the reference's topology prep:
  - `laplacian`, `rescale_L`  : lib/mesh_sampling.py:10-38 (same names, same arithmetic in fp32)
  - `load_graph_mtx`          : lib/load_data.py:7-32 (same return convention) -- reads the pickle-free
                                copy of data/transform_matrices/** made by tools/pack_topology.py
The reference turns every scipy matrix into a tf.SparseTensor and runs one SpMM per Chebyshev order and
per pool/unpool (lib/models.py:74-96,141-149).  Here the operators are constants, so they are composed
offline:  op_k = D . T_k(L~) . U  -- one sparse "row-gather" per polynomial order with pooling (row
selection) and unpooling (3-tap barycentric) folded in -- and packed as ELL tables for the CUDA kernels.
"""
import os

import numpy as np
import scipy.sparse as sp

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "smpl_topology.npz")
_cache = {}


def laplacian(W, normalized=True):
    """Normalised graph Laplacian I - D^-1/2 W D^-1/2 (lib/mesh_sampling.py:10-29)."""
    W = sp.csr_matrix(W)
    d = np.asarray(W.sum(axis=0)).reshape(-1).astype(W.dtype)
    if not normalized:
        return sp.csr_matrix(sp.diags(d, 0) - W)
    d = d + np.spacing(np.array(0, W.dtype))
    d = (1 / np.sqrt(d)).astype(W.dtype)
    Dm = sp.diags(d, 0)
    I = sp.identity(d.size, dtype=W.dtype)
    return sp.csr_matrix(I - Dm * W * Dm)


def rescale_L(L, lmax=2):
    """L/(lmax/2) - I on a copy (lib/mesh_sampling.py:31-38; chebyshev5 copies first, models.py:74)."""
    L = sp.csr_matrix(L, copy=True)
    I = sp.identity(L.shape[0], format="csr", dtype=L.dtype)
    L /= lmax / 2        # in place, as the reference does: keeps fp32 (scipy's out-of-place "/" upcasts to fp64)
    L -= I
    return sp.csr_matrix(L)


def _npz():
    if "npz" not in _cache:
        if not os.path.exists(_DATA):
            raise FileNotFoundError("%s missing: run tools/pack_topology.py in the build container" % _DATA)
        _cache["npz"] = np.load(_DATA)
    return _cache["npz"]


def _mats(kind, name, dtype):
    z = _npz()
    out = []
    for i in range(int(z["%s.%s.count" % (kind, name)])):
        k = "%s.%s.%d" % (kind, name, i)
        m = sp.csr_matrix((z[k + ".data"], z[k + ".indices"], z[k + ".indptr"]), shape=tuple(z[k + ".shape"]))
        out.append(m.astype(dtype))
    return out


def load_graph_mtx(project_dir=None, load_for_demo=False):
    """Same contract as lib/load_data.py:7-32: returns L_ds2, D_ds2, U_ds2 or, with load_for_demo,
    L, D, U, p, L_ds2, D_ds2, U_ds2 (all fp32; L = normalised Laplacians of the adjacency fixtures).
    `project_dir` is accepted for signature compatibility and ignored (fixtures live inside the package)."""
    A_ds2, D_ds2, U_ds2 = (_mats("ds2", n, np.float32) for n in "ADU")
    L_ds2 = [laplacian(a, normalized=True) for a in A_ds2]
    if not load_for_demo:
        return L_ds2, D_ds2, U_ds2
    A, D, U = (_mats("for_demo", n, np.float32) for n in "ADU")
    p = [a.shape[0] for a in A]
    L = [laplacian(a, normalized=True) for a in A]
    return L, D, U, p, L_ds2, D_ds2, U_ds2


def smpl_edges():
    """[20664, 2] int32 vertex pairs (data/edges_smpl.npy of the reference = upper triangle of A[0])."""
    return _npz()["edges"]


def trainset_stats():
    z = _npz()
    return z["stats.mean"], z["stats.std"]


def clothing_verts_idx():
    return _npz()["clothing_verts_idx"]


# ---------------------------------------------------------------------------------------------------
# operator algebra
# ---------------------------------------------------------------------------------------------------
def is_identity(S, tol=1e-9):
    S = sp.csr_matrix(S)
    if S.shape[0] != S.shape[1]:
        return False
    d = S - sp.identity(S.shape[0], dtype=S.dtype, format="csr")
    return d.nnz == 0 or float(np.abs(d.data).max()) <= tol


def cheb_polynomials(L, K):
    """[T_0(L~) .. T_{K-1}(L~)] as float64 CSR (sparse; intended for small K)."""
    Lt = rescale_L(sp.csr_matrix(L), lmax=2).astype(np.float64)
    Lt.eliminate_zeros()
    M = Lt.shape[0]
    T = [sp.identity(M, format="csr", dtype=np.float64)]
    if K > 1:
        T.append(Lt)
    for _ in range(2, K):
        T.append(sp.csr_matrix(2 * Lt @ T[-1] - T[-2]))
    return T


def compose(D, T, U):
    """D . T . U with None = identity; float64 CSR, explicit zeros removed."""
    m = sp.csr_matrix(T, dtype=np.float64)
    if U is not None:
        m = m @ sp.csr_matrix(U, dtype=np.float64)
    if D is not None:
        m = sp.csr_matrix(D, dtype=np.float64) @ m
    m = sp.csr_matrix(m)
    m.sum_duplicates()
    m.eliminate_zeros()
    m.sort_indices()
    return m


def to_ell(m):
    """CSR -> (idx int32 [rows, width], w fp32 [rows, width]); empty slots idx=-1, w=0, left-packed."""
    m = sp.csr_matrix(m)
    m.sort_indices()
    rows = m.shape[0]
    counts = np.diff(m.indptr)
    width = max(int(counts.max()) if rows else 1, 1)
    idx = np.full((rows, width), -1, np.int32)
    w = np.zeros((rows, width), np.float32)
    slot = np.arange(m.nnz) - np.repeat(m.indptr[:-1], counts)
    rr = np.repeat(np.arange(rows), counts)
    idx[rr, slot] = m.indices
    w[rr, slot] = m.data.astype(np.float32)
    return idx, w


def adjacency_ell(L):
    """Neighbour table of a level (off-diagonal pattern of its Laplacian), for the edge loss."""
    A = sp.csr_matrix(L, copy=True)
    A.setdiag(0)
    A.eliminate_zeros()
    A.data[:] = 1.0
    return to_ell(A)
