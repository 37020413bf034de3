#!/usr/bin/env python
"""Time the stand-alone gather (cape_resample) for the Laplacian term of each level, both vertex orders."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np, torch, scipy.sparse as sp
from cape_b200 import engine as E, topology as T

L, D, U, p, L2, D2, U2 = T.load_graph_mtx(load_for_demo=True)
tp = E.Topology(0)
orders = T.level_orders(L[0], D[:8])
N = 64
for lvl, F in ((0, 64), (0, 32), (2, 128), (4, 256), (6, 512), (7, 512)):
    Lt = T.cheb_polynomials(L[lvl], 2)[1]
    M = Lt.shape[0]
    x = torch.randn(N, M, F, device="cuda")
    y = torch.empty(N, M, F, device="cuda")
    line = "level %d M=%d F=%d:" % (lvl, M, F)
    for name, m in (("smpl", Lt), ("patch", T.permute(Lt, orders[lvl], orders[lvl]))):
        op = tp.add_operator(sp.csr_matrix(m))
        for _ in range(3):
            E.resample(tp, op, x, y, N, M, M, F)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            E.resample(tp, op, x, y, N, M, M, F)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        line += "  %s %.1fus (%.0f GB/s in+out)" % (name, us, 2 * N * M * F * 4 / us / 1e3)
    print(line, flush=True)
