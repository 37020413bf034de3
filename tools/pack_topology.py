#!/usr/bin/env python
"""One-off converter: reference operator fixtures -> pickle-free .npz inside this repo.

The reference ships its fixed mesh hierarchy as pickled scipy CSC matrices
(/root/reference/data/transform_matrices/{for_demo,ds2}/{A,D,U}.npy, loaded at
lib/load_data.py:7-32 with encoding='latin1').  The GPU box has no /root/reference, so the
operators are re-stored here as plain CSR arrays (indptr/indices/data/shape), loss-free.
The SMPL edge table (data/edges_smpl.npy, used by lib/losses.py:9-25) is the upper triangle
of A[0]; it is derived, checked against the reference file, and stored too.  The per-vertex
normalisation statistics (data/demo_data/trainset_stats.npz, demos.py:155) are stored for the
inference API.  The SMPL template itself is NOT copied: it cancels in the edge loss.

Run (only in the build container, where /root/reference exists):
    python tools/pack_topology.py
"""
import os
import sys
import numpy as np
import scipy.sparse as sp

REF = os.environ.get("CAPE_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cape_b200", "data", "smpl_topology.npz")


def _load(kind, name):
    path = os.path.join(REF, "data", "transform_matrices", kind, name + ".npy")
    return list(np.load(path, encoding="latin1", allow_pickle=True))


def main():
    out = {}
    for kind in ("for_demo", "ds2"):
        for name in ("A", "D", "U"):
            mats = _load(kind, name)
            out[f"{kind}.{name}.count"] = np.int64(len(mats))
            for i, m in enumerate(mats):
                m = sp.csr_matrix(m)
                m.sort_indices()
                key = f"{kind}.{name}.{i}"
                out[key + ".indptr"] = m.indptr.astype(np.int32)
                out[key + ".indices"] = m.indices.astype(np.int32)
                out[key + ".data"] = m.data  # dtype kept (for_demo fp32, ds2 fp64)
                out[key + ".shape"] = np.asarray(m.shape, np.int64)
    a0 = sp.coo_matrix(_load("for_demo", "A")[0])
    keep = a0.row < a0.col
    edges = np.stack([a0.row[keep], a0.col[keep]], 1).astype(np.int32)
    edges = edges[np.lexsort((edges[:, 1], edges[:, 0]))]
    ref_edges = np.load(os.path.join(REF, "data", "edges_smpl.npy"))
    assert set(map(tuple, edges.tolist())) == set(map(tuple, np.sort(ref_edges, 1).tolist()))
    out["edges"] = edges
    st = np.load(os.path.join(REF, "data", "demo_data", "trainset_stats.npz"))
    out["stats.mean"] = st["mean"].astype(np.float32)
    out["stats.std"] = st["std"].astype(np.float32)
    out["clothing_verts_idx"] = np.load(os.path.join(REF, "data", "clothing_verts_idx.npy")).astype(np.int32)
    np.savez_compressed(OUT, **out)
    print("wrote", os.path.abspath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
