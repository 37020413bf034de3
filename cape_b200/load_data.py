"""Dataset container and device-resident input pipeline.

`BodyData` keeps the reference's interface (lib/load_data.py:35-127): packed `.npy` files (or arrays) in, the last `nVal`
training examples split off as validation set, per-vertex mean/std of the training split, normalised float32 arrays
`vertices_{train,val,test}`, `cond1_*` (pose, reduced to the 14 clothing-related joints when full poses are given),
`cond2_*` (clothing type).  No psbody: the reference mesh is read by a small OBJ reader.

`DeviceDataset` is what replaces the reference's per-step numpy fancy-indexing + feed_dict (lib/models.py:877-903): the
normalised training split lives in HBM once (the CAPE dataset is ~3.5 GB of 180 GB), and a step's batch is assembled
by a gather kernel from indices -- per step only 2 N int32 indices and the N x nz noise cross PCIe instead of three
[N, 6890, 3] tensors.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .demos import filter_cloth_pose, read_obj


def _load(x):
    return np.load(x) if isinstance(x, str) else np.asarray(x)


class BodyData(object):
    def __init__(self, nVal, train_mesh_fn, train_cond1_fn, test_mesh_fn, test_cond1_fn, reference_mesh_file=None,
                 train_cond2_fn=None, test_cond2_fn=None):
        self.nVal = nVal
        self.train_mesh_fn, self.train_cond1_fn, self.train_cond2_fn = train_mesh_fn, train_cond1_fn, train_cond2_fn
        self.test_mesh_fn, self.test_cond1_fn, self.test_cond2_fn = test_mesh_fn, test_cond1_fn, test_cond2_fn
        self.vertices_train = self.cond1_train = self.vertices_val = self.cond1_val = None
        self.vertices_test = self.cond1_test = None
        self.N = self.n_vertex = None
        self.load()
        self.reference_mesh = read_obj(reference_mesh_file) if reference_mesh_file else None     # (vertices, faces)
        self.mean = np.mean(self.vertices_train, axis=0)
        self.std = np.std(self.vertices_train, axis=0)
        self.normalize()
        self.change_dtype()

    def load(self):
        vertices_train = np.array(_load(self.train_mesh_fn), dtype=np.float64)       # a copy: normalised in place below
        self.vertices_train = vertices_train[:-self.nVal]
        self.vertices_val = vertices_train[-self.nVal:]
        cond1_train = _load(self.train_cond1_fn)
        if len(cond1_train.shape) > 2:                       # pose param not flattened
            cond1_train = cond1_train.reshape(len(cond1_train), -1)
        self.cond1_train, self.cond1_val = cond1_train[:-self.nVal], cond1_train[-self.nVal:]
        if self.train_cond2_fn is not None:
            cond2_train = _load(self.train_cond2_fn)
            self.cond2_train, self.cond2_val = cond2_train[:-self.nVal], cond2_train[-self.nVal:]
        self.n_vertex = self.vertices_train.shape[1]
        self.vertices_test = np.array(_load(self.test_mesh_fn), dtype=np.float64)
        self.cond1_test = _load(self.test_cond1_fn)
        if self.test_cond2_fn is not None:
            self.cond2_test = _load(self.test_cond2_fn)
        if len(self.cond1_test.shape) > 2:
            self.cond1_test = self.cond1_test.reshape(len(self.cond1_test), -1)
        # remove the pose parameters of joints irrelevant to clothing, keep the full ones for re-posing (:93-98)
        if self.cond1_test.shape[-1] % 14 != 0:
            self.cond1_test_full, self.cond1_train_full, self.cond1_val_full = self.cond1_test, self.cond1_train, self.cond1_val
            self.cond1_train, self.cond1_val, self.cond1_test = list(map(filter_cloth_pose, [self.cond1_train, self.cond1_val,
                                                                                              self.cond1_test]))
        print("Data loaded, {} train, {} val, {} test examples.\n".format(len(self.vertices_train), len(self.vertices_val),
                                                                          len(self.vertices_test)))

    def normalize(self):
        for a in (self.vertices_train, self.vertices_val, self.vertices_test):
            a -= self.mean
            a /= self.std
        print("Vertices normalized.\n")

    def change_dtype(self):
        for k in ("vertices_train", "vertices_val", "vertices_test", "cond1_train", "cond1_val", "cond1_test"):
            setattr(self, k, getattr(self, k).astype("float32"))
        if self.train_cond2_fn is not None:
            for k in ("cond2_train", "cond2_val", "cond2_test"):
                setattr(self, k, getattr(self, k).astype("float32"))

    def vec2mesh(self, vec):
        """(vertices, faces) of a de-normalised prediction (the reference returns a psbody Mesh)."""
        vec = np.asarray(vec).reshape((self.n_vertex, 3)) * self.std + self.mean
        return vec, (self.reference_mesh[1] if self.reference_mesh else None)

    def get_normalized_meshes(self, mesh_paths):
        return np.array([(read_obj(p)[0] - self.mean) / self.std for p in mesh_paths])


class DeviceDataset:
    """Training split resident on the GPU + batch assembly by index (cape_gather_rows)."""

    def __init__(self, vertices, cond1, cond2, device):
        f = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(device)
        self.vertices, self.cond1, self.cond2 = f(vertices), f(cond1), f(cond2)
        self.n = int(self.vertices.shape[0])
        self.device = self.vertices.device
        self.lib = _lib.load()
        self._idx_host = [torch.zeros(0, dtype=torch.int32).pin_memory() for _ in range(4)]     # ring of pinned index buffers
        self._slot = 0

    @classmethod
    def from_body_data(cls, data, device, split="train"):
        return cls(getattr(data, "vertices_" + split), getattr(data, "cond1_" + split), getattr(data, "cond2_" + split),
                   device)

    def nbytes(self):
        return 4 * (self.vertices.numel() + self.cond1.numel() + self.cond2.numel())

    def _upload(self, idx):
        idx = np.asarray(idx, np.int32).reshape(-1)
        slot = self._slot
        self._slot = (slot + 1) % len(self._idx_host)
        if self._idx_host[slot].numel() < idx.size:
            self._idx_host[slot] = torch.zeros(idx.size, dtype=torch.int32).pin_memory()
        h = self._idx_host[slot][: idx.size]
        h.copy_(torch.from_numpy(idx))
        return h.to(self.device, non_blocking=True)

    def _gather(self, src, idx_dev, dst):
        rowf = src[0].numel()
        assert dst.is_contiguous() and dst.shape[0] == idx_dev.numel() and dst[0].numel() == rowf
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(self.lib.cape_gather_rows(C.c_void_p(src.data_ptr()), rowf, self.n, C.c_void_p(idx_dev.data_ptr()),
                                             idx_dev.numel(), C.c_void_p(dst.data_ptr()), st))

    def stage(self, net, idx_g, idx_d=None, eps=None):
        """Assemble the generator (and discriminator) batch of a step in the network's input buffers: one small H2D copy
        of the indices (+ the noise), three gathers per batch on the compute stream."""
        N = net.N
        both = np.concatenate([idx_g, idx_d]) if idx_d is not None else np.asarray(idx_g)
        dev = self._upload(both)
        ig = dev[:N]
        self._gather(self.vertices, ig, net.in_x)
        self._gather(self.cond1, ig, net.in_cond[N:])
        self._gather(self.cond2, ig, net.in_cond2[N:])
        if idx_d is not None:
            idd = dev[N:]
            self._gather(self.vertices, idd, net.xcat[:N])
            self._gather(self.cond1, idd, net.in_cond[:N])
            self._gather(self.cond2, idd, net.in_cond2[:N])
        if eps is not None:
            net.in_eps.copy_(eps if torch.is_tensor(eps) else torch.from_numpy(np.ascontiguousarray(eps, np.float32)),
                             non_blocking=True)
