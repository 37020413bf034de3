#!/usr/bin/env python
"""Debug runner: where does the CUDA backward leave the fp64 truth?  Compares hidden encoder activations and
their gradients (net.enc_act / net.g_enc / net.g_enc_red) with the float64 oracle, with the reference's own
initialisers (fc_scale=1.0), tensor cores on and off.  Not a test; tests/test_gpu_parity.py asserts on the
same quantities through parity.train_step_truth."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    import numpy as np
    import torch
    import parity
    from oracle import cape_oracle as O
    from cape_b200 import topology as T
    from cape_b200.network import CapeNetwork
    from cape_b200.params import NZ64_AFFINE, param_specs
    from cape_b200.synthetic import make_batch
    L, D, U, p, L_d, D_d, U_d = T.load_graph_mtx(load_for_demo=True)
    h = dict(L=L, D=D, U=U, p=p, L_d=L_d, D_d=D_d)
    cfg = dict(NZ64_AFFINE, decay_steps=10)
    N = 2
    specs = param_specs(cfg, [l.shape[0] for l in L], [l.shape[0] for l in L_d])
    params = parity.calibrated_params(specs, 123, float(os.environ.get("FC_SCALE", "1.0")))
    batch = make_batch(N, cfg["nz"], seed=123)
    tb = {k: torch.from_numpy(v) for k, v in batch.items()}
    for tc_on in (True, False):
        prev = parity.set_tensor_cores(tc_on)
        net = CapeNetwork(L, D, U, L_d, D_d, cfg, N, params=params)
        net.set_inputs(tb["x_g"], tb["cond_g"], tb["cond2_g"], tb["eps"], tb["x_d"], tb["cond_d"], tb["cond2_d"])
        net.train_step(step=100, update=False)
        torch.cuda.synchronize()
        parity.set_tensor_cores(prev)
        o = O.Oracle(L, D, U, L_d, D_d, cfg, dtype=torch.float64)
        o.masks, o.mask_rows = parity.cuda_masks(net, h, N)
        o.keep = {}
        P = {k: torch.from_numpy(v).double().requires_grad_(True) for k, v in params.items()}
        ob = {k: v.double() for k, v in tb.items()}
        y, y2 = o.cond_embeddings(ob["cond_g"], ob["cond2_g"], P, tag="_g")
        yd, y2d = o.cond_embeddings(ob["cond_d"], ob["cond2_d"], P, tag="_d")
        x_hat, zm, zl = o.generator(ob["x_g"], y, y2, ob["eps"], P)
        d_real = o.discriminator(ob["x_d"], yd, y2d, P, tag="_real")
        d_fake = o.discriminator(x_hat, y, y2, P, tag="_fake")
        Ls = o.losses(x_hat, ob["gt"], zm, zl, d_real, d_fake, P, T.smpl_edges())
        o.keep["x_hat"] = x_hat
        names = list(o.keep)
        gr = torch.autograd.grad(Ls["loss_g"], [o.keep[k] for k in names] + [zm, zl], allow_unused=True, retain_graph=True)
        G = dict(zip(names + ["z_mean", "z_logvar"], gr))
        print("==== tensor cores %s ====" % ("on" if tc_on else "off"))
        print("  z_mean      fwd %.2e   z_logvar fwd %.2e" % (parity.rel(net.z_mean.cpu().numpy(), zm.detach().numpy()),
                                                          parity.rel(net.z_logvar.cpu().numpy(), zl.detach().numpy())))
        print("  g_mean %.2e  g_logvar %.2e" % (parity.rel(net.g_mean.cpu().numpy(), G["z_mean"].numpy()),
                                              parity.rel(net.g_logvar.cpu().numpy(), G["z_logvar"].numpy())))
        print("  enc_red     fwd %.2e   grad %.2e" % (parity.rel(net.enc_red.cpu().numpy(), o.keep["enc_red"].detach().numpy()),
                                                     parity.rel(net.g_enc_red.cpu().numpy(), G["enc_red"].numpy())))
        print("  z_total     fwd %.2e   (max |z| %.3e)" % (parity.rel(net.z_total.cpu().numpy(), o.keep["z_total"].detach().numpy()),
                                                      float(o.keep["z_total"].abs().max())))
        print("  x_hat       fwd %.2e   (max |x_hat| %.3e)   d_xhat %.2e (max %.3e)" % (
            parity.rel(net.x_hat.cpu().numpy(), x_hat.detach().numpy()), float(x_hat.abs().max()),
            parity.rel(net.d_xhat.cpu().numpy(), G["x_hat"].numpy()), float(G["x_hat"].abs().max())))
        for nm, L_ in (("recon", Ls["recon"] * cfg["lambda_recon"]), ("edge", Ls["edge"] * cfg["lambda_edge"]),
                       ("gan_g", Ls["gan_g"] * cfg["lambda_gan"])):
            gg = torch.autograd.grad(L_, x_hat, retain_graph=True)[0]
            err = (net.d_xhat.cpu().double() - G["x_hat"]).abs()
            wi = np.unravel_index(int(err.argmax()), err.shape)
            print("      d %-6s / d x_hat: max %.3e   at the worst element %s: %.4e  (cuda total %.4e, oracle total %.4e, "
                  "x_hat - x there: cuda %.4e oracle %.4e)" % (nm, float(gg.abs().max()), wi, float(gg[wi]),
                                                               float(net.d_xhat[wi]), float(G["x_hat"][wi]),
                                                               float(net.x_hat[wi] - net.in_x[wi]),
                                                               float(x_hat[wi] - ob["gt"][wi])))
        a = o.keep["dec_fc"].detach()
        slope = torch.where(a > 0, torch.ones_like(a), torch.full_like(a, 0.2))
        print("  dec_fc      fwd %.2e   grad(pre) %.2e   grad(z_total) %.2e" % (
            parity.rel(net.dec_fc.cpu().numpy(), a.numpy()),
            parity.rel(net.g_dec_fc_t.cpu().numpy(), (G["dec_fc"] * slope).numpy()),
            parity.rel(net.g_z.cpu().numpy(), G["z_total"][:, :net.nz].numpy())))
        for i in range(1, 9):
            a = o.keep["dec_act%d" % i].detach()
            print("  dec_act%d    fwd %.2e   (max %.3e)" % (i, parity.rel(net.dec_act[i - 1].cpu().numpy(), a.numpy()),
                                                          float(a.abs().max())))
        for i in range(8, 0, -1):
            a = o.keep["enc_act%d" % i].detach()
            # net.g_enc[i-1] is the gradient w.r.t. the PRE-activation (slope already applied), pooled rows
            slope = torch.where(a > 0, torch.ones_like(a), torch.full_like(a, 0.2))
            gpre = (G["enc_act%d" % i] * slope).numpy()
            got = net.g_enc[i - 1].cpu().numpy()
            err = np.abs(got - gpre)
            print("  enc_act%d    fwd %.2e   grad(pre) %.2e   (max |g| %.3e, worst at %s)" % (
                i, parity.rel(net.enc_act[i - 1].cpu().numpy(), a.numpy()), parity.rel(got, gpre), np.abs(gpre).max(),
                np.unravel_index(err.argmax(), err.shape)))


if __name__ == "__main__":
    main()
