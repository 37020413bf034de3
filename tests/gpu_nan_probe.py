#!/usr/bin/env python
"""Debug runner: forward of the nz64 model at batch N with the reference initialisers; reports the first buffer with
non-finite values and compares every decoder activation between layer forms (CAPE_MODES / CAPE_FWD_MODE)."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def run(N, env):
    import torch
    import parity
    from cape_b200 import topology as T
    from cape_b200.network import CapeNetwork
    from cape_b200.params import NZ64_AFFINE, param_specs
    from cape_b200.synthetic import make_batch
    for k in ("CAPE_FWD_MODE", "CAPE_DX_MODE", "CAPE_MODES"):
        os.environ.pop(k, None)
    os.environ.update(env)
    L, D, U, p, L_d, D_d, _ = T.load_graph_mtx(load_for_demo=True)
    cfg = dict(NZ64_AFFINE, decay_steps=10)
    specs = param_specs(cfg, [l.shape[0] for l in L], [l.shape[0] for l in L_d])
    params = parity.calibrated_params(specs, 123, 1.0)
    net = CapeNetwork(L, D, U, L_d, D_d, cfg, N, params=params)
    tb = {k: torch.from_numpy(v) for k, v in make_batch(N, cfg["nz"], seed=123).items()}
    net.set_inputs(tb["x_g"], tb["cond_g"], tb["cond2_g"], tb["eps"], tb["x_d"], tb["cond_d"], tb["cond2_d"])
    net.train_step(step=100, update=False)
    torch.cuda.synchronize()
    out = {"z_total": net.z_total, "dec_fc": net.dec_fc, "dec_h0": net.dec_h0}
    for i, a in enumerate(net.enc_act):
        out["enc_act%d" % (i + 1)] = a
    for i, a in enumerate(net.dec_act):
        out["dec_act%d" % (i + 1)] = a
    out["x_hat"] = net.x_hat
    for i, a in enumerate(net.disc_act):
        out["disc_act%d" % (i + 1)] = a
    out["logits"] = net.logits
    for i, a in enumerate(net.g_dec):
        out["g_dec%d" % (i + 1)] = a
    for i, a in enumerate(net.g_enc):
        out["g_enc%d" % (i + 1)] = a
    out["grad_G"] = net.PG.grad
    out["grad_D"] = net.PD.grad
    return {k: v.clone() for k, v in out.items()}, [(l.name, l.fwd_mode, l.dx_mode) for l in net.all_layers()]


def main():
    import torch
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    a, modes = run(N, {})
    print("modes:", [m for m in modes if m[1] != "fused" or m[2] != "fused"])
    b, _ = run(N, {"CAPE_FWD_MODE": "fused", "CAPE_DX_MODE": "fused"})
    for k in a:
        fa, fb = bool(torch.isfinite(a[k]).all()), bool(torch.isfinite(b[k]).all())
        d = float((a[k] - b[k]).abs().max() / b[k].abs().max().clamp_min(1e-30)) if fa and fb else float("nan")
        print("%-12s finite(default forms) %-5s finite(all fused) %-5s  max-rel diff %.2e   max|.| %.3e" % (
            k, fa, fb, d, float(b[k].abs().max()) if fb else float("nan")))


if __name__ == "__main__":
    main()
