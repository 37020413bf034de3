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




def where():
    """Locate the non-finite entries of the first decoder block's output (batch 64, default forms)."""
    import torch
    import parity
    N = 64
    a, _ = run(N, {})
    d = a["dec_act1"]
    bad = ~torch.isfinite(d)
    print("dec_act1 shape", tuple(d.shape), "non-finite", int(bad.sum()))
    print("per sample:", bad.flatten(1).sum(1).tolist())
    rows = bad.any(2).nonzero()
    print("first bad (n, row):", rows[:10].tolist(), "last:", rows[-5:].tolist())
    cols = bad.any(0).any(0).nonzero().flatten()
    print("bad columns:", cols[:20].tolist(), "... count", cols.numel())
    flat = bad.any(2).flatten().nonzero().flatten()
    print("bad flat rows // 128 (tiles):", sorted(set((flat // 128).tolist()))[:40])
    prev = parity.set_tensor_cores(False)
    b, _ = run(N, {})
    parity.set_tensor_cores(prev)
    print("tensor cores off: dec_act1 finite", bool(torch.isfinite(b["dec_act1"]).all()), " x_hat finite",
          bool(torch.isfinite(b["x_hat"]).all()))


if len(sys.argv) > 2 and sys.argv[2] == "where":
    where()


def knobs():
    """Forward only (no backward), batch 64: is the first decoder block's output finite under the experiment knobs?"""
    import torch
    import parity
    from cape_b200 import _lib, topology as T
    from cape_b200.network import CapeNetwork
    from cape_b200.params import NZ64_AFFINE, param_specs
    from cape_b200.synthetic import make_batch
    lib = _lib.load()
    L, D, U, p, L_d, D_d, _ = T.load_graph_mtx(load_for_demo=True)
    cfg = dict(NZ64_AFFINE, decay_steps=10)
    specs = param_specs(cfg, [l.shape[0] for l in L], [l.shape[0] for l in L_d])
    params = parity.calibrated_params(specs, 123, 1.0)
    N = 64
    tb = {k: torch.from_numpy(v) for k, v in make_batch(N, cfg["nz"], seed=123).items()}
    for label, kn in (("default", {}), ("knob6 (identity tiles by producers)", {6: 1}), ("knob4 (weights by producers)", {4: 1}),
                      ("knob8 (no gemm_tc)", {8: 1}), ("knob4+6", {4: 1, 6: 1})):
        for k, v in kn.items():
            lib.cape_set_tuning(k, v)
        net = CapeNetwork(L, D, U, L_d, D_d, cfg, N, params=params)
        net.set_inputs(tb["x_g"], tb["cond_g"], tb["cond2_g"], tb["eps"])
        net.forward_generator()
        torch.cuda.synchronize()
        bad = [int((~torch.isfinite(a)).sum()) for a in net.dec_act]
        print("%-40s non-finite per decoder block %s   dec_rg1 %d  x_hat %d" % (
            label, bad, int((~torch.isfinite(net.dec_rg[0])).sum()), int((~torch.isfinite(net.x_hat)).sum())))
        # run the first block again on its own: is it reproducible in isolation?
        net.dec[0].fwd(net.dec_h0, net.ycat_g, net.dec_act[0], out2=net.dec_rg[0])
        torch.cuda.synchronize()
        print("%-40s   block 1 alone: %d" % ("", int((~torch.isfinite(net.dec_act[0])).sum())))
        for k in kn:
            lib.cape_set_tuning(k, 0)
        del net


if len(sys.argv) > 2 and sys.argv[2] == "knobs":
    knobs()
if __name__ == "__main__" and len(sys.argv) <= 2:
    main()


def isolate():
    """The first decoder block alone (wide fused kernel, two accumulators) at batch 64 with doctored inputs."""
    import torch
    import parity
    from cape_b200 import _lib, topology as T
    from cape_b200.network import CapeNetwork
    from cape_b200.params import NZ64_AFFINE, param_specs
    L, D, U, p, L_d, D_d, _ = T.load_graph_mtx(load_for_demo=True)
    cfg = dict(NZ64_AFFINE, decay_steps=10)
    specs = param_specs(cfg, [l.shape[0] for l in L], [l.shape[0] for l in L_d])
    params = parity.calibrated_params(specs, 123, 1.0)
    for N in (24, 64):
        net = CapeNetwork(L, D, U, L_d, D_d, cfg, N, params=params)
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(N, 862, 512, device="cuda", generator=g)
        y = torch.randn(N, 64, device="cuda", generator=g)
        lay = net.dec[0]

        def go(label, xx, yy):
            lay.fwd(xx, yy, net.dec_act[0], out2=net.dec_rg[0])
            torch.cuda.synchronize()
            o = net.dec_act[0]
            bad = ~torch.isfinite(o)
            rows = bad.any(2).flatten().nonzero().flatten()
            print("N=%d %-26s non-finite %7d  rows%%128 %s  cols %s  first values %s" % (
                N, label, int(bad.sum()), sorted(set((rows % 128).tolist()))[:4],
                (bad.any(0).any(0).nonzero().flatten()[[0, -1]].tolist() if bad.any() else []),
                o[0, 0, :3].tolist()))

        go("random x, random cond", x, y)
        go("random x, zero cond", x, torch.zeros_like(y))
        go("zero x, random cond", torch.zeros_like(x), y)
        go("ones x, zero cond", torch.ones_like(x), torch.zeros_like(y))
        lay.Wt[2].zero_(); lay.Wt_lo[2].zero_()
        go("affine K-major copy zeroed", x, y)
        del net


if len(sys.argv) > 2 and sys.argv[2] == "isolate":
    isolate()
