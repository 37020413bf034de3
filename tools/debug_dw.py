#!/usr/bin/env python
"""Probe the tcgen05 weight-gradient kernel with structured inputs (debug aid)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cape_b200 import engine as E, ops, _lib

tp = ops.topology_for(torch.device("cuda", 0))
lib = _lib.load()

def run(x, g, F, ncols, tc):
    lib.cape_set_tensor_cores(1 if tc else 0)
    N, rows = x.shape[0], x.shape[1]
    dw = torch.full((F, ncols), -7.0, device="cuda")
    E.cheb_dw(tp, N, rows, ncols, x, -1, F, rows, x.shape[2], g, dw, ncols)
    torch.cuda.synchronize()
    return dw.cpu().numpy()

def report(tag, a, b):
    err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
    print("%-40s rel=%.3e  tc[min,max]=[%.3g,%.3g] ref[min,max]=[%.3g,%.3g] nnz_tc=%d nnz_ref=%d" % (
        tag, err, a.min(), a.max(), b.min(), b.max(), (a != 0).sum(), (b != 0).sum()))
    return err

N, rows = 2, 4096
for F, ncols in ((128, 128), (64, 64), (256, 512), (128, 32)):
    print("== F=%d ncols=%d" % (F, ncols))
    # 1) one-hot channels
    x = torch.zeros(N, rows, F, device="cuda"); g = torch.zeros(N, rows, ncols, device="cuda")
    f0, c0 = 5 % F, 9 % ncols
    x[:, :, f0] = 1.0; g[:, :, c0] = 1.0
    a, b = run(x, g, F, ncols, True), run(x, g, F, ncols, False)
    report("one-hot f0=%d c0=%d" % (f0, c0), a, b)
    print("   tc argmax", np.unravel_index(np.argmax(np.abs(a)), a.shape), "val", a.flat[np.argmax(np.abs(a))],
          " ref argmax", np.unravel_index(np.argmax(np.abs(b)), b.shape), "val", b.flat[np.argmax(np.abs(b))])
    nz = np.argwhere(np.abs(a) > 1e-3)[:12]
    print("   tc nonzeros (first 12):", [(int(i), int(j), float(a[i, j])) for i, j in nz])
    # 2) f-ramp x c-ramp with constant rows: dW[f,c] = rows_total * (f+1) * (c+1) * 1e-4
    x = (torch.arange(F, device="cuda").float() + 1).view(1, 1, F).expand(N, rows, F).contiguous() * 1e-2
    g = (torch.arange(ncols, device="cuda").float() + 1).view(1, 1, ncols).expand(N, rows, ncols).contiguous() * 1e-2
    a, b = run(x, g, F, ncols, True), run(x, g, F, ncols, False)
    report("ramps", a, b)
    print("   tc[0:3,0:4]=%s\n   ref[0:3,0:4]=%s" % (a[:3, :4].round(3).tolist(), b[:3, :4].round(3).tolist()))
    print("   tc[:,0][:8]=%s  ref=%s" % (a[:8, 0].round(3).tolist(), b[:8, 0].round(3).tolist()))
    # 3) only one row active
    x = torch.zeros(N, rows, F, device="cuda"); g = torch.zeros(N, rows, ncols, device="cuda")
    x[0, 37, :] = torch.arange(F, device="cuda").float() + 1
    g[0, 37, :] = 1.0
    a, b = run(x, g, F, ncols, True), run(x, g, F, ncols, False)
    report("single row 37", a, b)
    print("   tc[:8,0]=%s ref[:8,0]=%s" % (a[:8, 0].round(3).tolist(), b[:8, 0].round(3).tolist()))
    # 4) random
    x = torch.randn(N, rows, F, device="cuda"); g = torch.randn(N, rows, ncols, device="cuda")
    a, b = run(x, g, F, ncols, True), run(x, g, F, ncols, False)
    report("random", a, b)
lib.cape_set_tensor_cores(1)
