"""Synthetic inputs of the shape the reference trains on (there is no dataset offline).

Offsets x ~ N(0,1) [N,6890,3] (the network sees per-vertex standardised data, lib/load_data.py:103-113);
pose condition [N,126] = 14 joint rotation matrices from axis-angle ~ N(0, 0.3^2) via Rodrigues
(lib/prep_data.py:76-77 + lib/utils.py:38-62); clothing condition [N,4] one-hot; eps ~ N(0,1) [N,nz].
"""
import numpy as np


def rodrigues(aa):
    """Axis-angle [..., 3] -> rotation matrices [..., 3, 3]."""
    aa = np.asarray(aa, np.float64)
    th = np.linalg.norm(aa, axis=-1, keepdims=True)
    k = aa / np.maximum(th, 1e-12)
    K = np.zeros(aa.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    s, c = np.sin(th)[..., None], np.cos(th)[..., None]
    return np.eye(3) + s * K + (1 - c) * (K @ K)


def make_batch(N, nz, seed=123, n_verts=6890, n_joints=14, n_clo=4):
    """Dict of float32 arrays: x_g, gt, cond_g, cond2_g, eps, x_d, cond_d, cond2_d."""
    rng = np.random.RandomState(seed)

    def conds():
        aa = rng.normal(0, 0.3, size=(N, n_joints, 3))
        c1 = rodrigues(aa).reshape(N, n_joints * 9).astype(np.float32)
        c2 = np.eye(n_clo, dtype=np.float32)[rng.randint(0, n_clo, size=N)]
        return c1, c2

    x_g = rng.normal(size=(N, n_verts, 3)).astype(np.float32)
    cond_g, cond2_g = conds()
    eps = rng.normal(size=(N, nz)).astype(np.float32)
    x_d = rng.normal(size=(N, n_verts, 3)).astype(np.float32)
    cond_d, cond2_d = conds()
    return dict(x_g=x_g, gt=x_g, cond_g=cond_g, cond2_g=cond2_g, eps=eps, x_d=x_d, cond_d=cond_d, cond2_d=cond2_d)
