#!/usr/bin/env python
"""Probe/time the dense weight-gradient kernels (TMA+tcgen05 vs gather-tcgen05 vs SIMT) on op = identity."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cape_b200 import engine as E, ops, _lib

tp = ops.topology_for(torch.device("cuda", 0))
tp.reserve_workspace(64 << 20)
lib = _lib.load()


def run(x, g, F, ncols, mode, reps=0):
    """mode: 'tma0' / 'tma1' (lo-part variants), 'tc' (gather kernel), 'simt'"""
    lib.cape_set_tensor_cores(0 if mode == "simt" else 1)
    lib.cape_set_tuning(1, 0 if mode.startswith("tma") else 1)
    lib.cape_set_tuning(2, 1 if mode == "tma1" else 0)
    lib.cape_set_tuning(3, 2 if mode == "tma128" else 0)
    N, rows = x.shape[0], x.shape[1]
    dw = torch.full((F, ncols), -7.0, device="cuda")
    E.cheb_dw(tp, N, rows, ncols, x, -1, F, rows, x.shape[2], g, dw, ncols)
    torch.cuda.synchronize()
    ms = None
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            E.cheb_dw(tp, N, rows, ncols, x, -1, F, rows, x.shape[2], g, dw, ncols)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    return dw.double().cpu().numpy(), ms


shapes = [(3, 6890, 96, 96), (2, 4096, 128, 128), (64, 862, 512, 512), (64, 1723, 256, 512), (64, 3445, 128, 256),
          (64, 6890, 128, 64), (64, 6890, 64, 64), (64, 6890, 64, 32), (128, 3445, 128, 128)]
for N, rows, F, ncols in shapes:
    torch.manual_seed(0)
    x = torch.randn(N, rows, F, device="cuda")
    g = torch.randn(N, rows, ncols, device="cuda")
    ref = (x.double().reshape(-1, F).t() @ g.double().reshape(-1, ncols)).cpu().numpy()
    scale = np.abs(ref).max()
    line = "N=%d rows=%d F=%d ncols=%d :" % (N, rows, F, ncols)
    for mode in ("tma0", "tma128", "tc"):
        try:
            got, ms = run(x, g, F, ncols, mode, reps=10)
            line += "  %s err=%.2e %.1fus" % (mode, np.abs(got - ref).max() / scale, ms * 1e3)
        except Exception as e:  # noqa
            line += "  %s FAILED(%s)" % (mode, str(e)[:60])
    print(line, flush=True)
    # structured: one active row, checks layout/swizzle exactly
    x = torch.zeros(N, rows, F, device="cuda"); g = torch.zeros(N, rows, ncols, device="cuda")
    x[0, 37, :] = torch.arange(F, device="cuda").float() + 1
    g[0, 37, :] = torch.arange(ncols, device="cuda").float() * 0.5 + 1
    ref = (x.double().reshape(-1, F).t() @ g.double().reshape(-1, ncols)).cpu().numpy()
    got, _ = run(x, g, F, ncols, "tma0")
    print("   single-row exactness: max abs err %.3e (ref max %.1f)" % (np.abs(got - ref).max(), np.abs(ref).max()))
lib.cape_set_tensor_cores(1); lib.cape_set_tuning(1, 0); lib.cape_set_tuning(2, 0); lib.cape_set_tuning(3, 0)
