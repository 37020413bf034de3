"""CPU oracle for CAPE's graph-conv hot path -- TEST INFRASTRUCTURE ONLY.

A torch-CPU (autograd) restatement of the reference's TF-1.13 graph, function by function, citing
/root/reference/lib/models.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg may import this file; the product path (cape_b200/) never does.

PINNING: the reference ships no tests or golden vectors and its arithmetic lives in
tensorflow-gpu==1.13.2 (requirements.txt:11), which cannot be installed here -- TensorFlow's kernels themselves are
therefore unpinned.  The reference's MODEL CODE is pinned: lib/models.py is executed unmodified on a TF-1 API shim
(oracle/tf1_shim.py, torch-CPU behind the ~70 symbols it calls) and its outputs for one full update of both model
families, the demo-phase graph and the op bodies are committed as golden vectors (tests/golden/make_ref_golden.py ->
ref_models_golden.npz); tests/test_reference_golden.py requires this file's functions to reproduce them (forward: bit
for bit; gradients / updates: 4e-7).  Also pinned: the reference's own importable host code (lib/mesh_sampling.py
laplacian/rescale_L; tests/golden/make_golden.py), an independent float64 dense-polynomial formulation of the Chebyshev
conv and a literal numpy/scipy transcription of the op bodies (oracle/np_ops.py).

TF-default semantics encoded here (TF-1.13 docs): tf.nn.leaky_relu alpha=0.2; tf.layers.dense y=xW+b;
tf.losses.* Reduction.MEAN over all elements; l2_regularizer(s)(w) = s*sum(w^2)/2; MomentumOptimizer
a <- m*a + g, w <- w - lr*a; clip_by_global_norm scale = clip/max(norm, clip);
exponential_decay(staircase=True) uses floor.
"""
import math
from collections import OrderedDict

import numpy as np
import scipy.sparse as sp
import torch


# --------------------------------------------------------------------------------------------------
# host-side operator prep (lib/mesh_sampling.py:10-38)
# --------------------------------------------------------------------------------------------------
def laplacian(W, normalized=True):
    """lib/mesh_sampling.py:10-29."""
    d = W.sum(axis=0)
    if not normalized:
        D = sp.diags(np.asarray(d).squeeze(), 0)
        L = D - W
    else:
        d = d + np.spacing(np.array(0, W.dtype))
        d = 1 / np.sqrt(d)
        D = sp.diags(np.asarray(d).squeeze(), 0)
        I = sp.identity(d.size, dtype=W.dtype)
        L = I - D * W * D
    return sp.csr_matrix(L)


def rescale_L(L, lmax=2):
    """lib/mesh_sampling.py:31-38 (on a copy, as chebyshev5 does at models.py:74)."""
    L = sp.csr_matrix(L, copy=True)
    M = L.shape[0]
    I = sp.identity(M, format="csr", dtype=L.dtype)
    L /= lmax / 2        # in place (mesh_sampling.py:36-37): keeps the fp32 dtype
    L -= I
    return sp.csr_matrix(L)


def _to_torch_sparse(m, dtype):
    m = sp.coo_matrix(m)
    idx = torch.from_numpy(np.vstack([m.row, m.col]).astype(np.int64))
    val = torch.from_numpy(m.data.astype(np.float64)).to(dtype)
    return torch.sparse_coo_tensor(idx, val, m.shape).coalesce()


class Oracle:
    """Functional mirror of base_model + CAPE (lib/models.py:13-832).

    params: dict name -> torch tensor (requires_grad as the caller wishes), names are the reference's TF
    variable names ('generator/encoder/encoder_conv1/weights', ...).
    """

    def __init__(self, L, D, U, L_d, D_d, cfg, dtype=torch.float32):
        self.cfg = cfg
        self.dtype = dtype
        self.F = list(cfg["F"])
        self.K = list(cfg["K"])
        self.Kd = cfg["Kd"]
        self.p = [int(l.shape[0]) for l in L]
        # models.py:74-79: rescaled Laplacian as a sparse tensor (rescale done per call there; hoisted here)
        self.Lt = [_to_torch_sparse(rescale_L(l, lmax=2), dtype) for l in L]
        self.Lt_d = [_to_torch_sparse(rescale_L(l, lmax=2), dtype) for l in L_d]
        self.Dm = [_to_torch_sparse(d, dtype) for d in D]
        self.Um = [_to_torch_sparse(u, dtype) for u in U]
        self.Dm_d = [_to_torch_sparse(d, dtype) for d in D_d]
        rd = cfg.get("reduce_dim", 64)
        self.reduce_rate = self.F[-1] // rd if rd > 0 else 1  # models.py:254-257
        self.reduce_dim = rd
        # Optional {site name: bool tensor}: branch decisions (pre-activation > 0) imposed on the (leaky-)ReLUs.
        # Two fp32 implementations round pre-activations differently, so ~1e-6 of them change sign; with the
        # masks of the implementation under test imposed, the gradient comparison is exact up to rounding
        # instead of being dominated by a handful of flipped units.  Entries may contain only some rows
        # (pooled sites): `mask_rows[name]` then lists the rows they refer to.
        self.masks = None
        self.mask_rows = {}
        # Optional dict: when set, every (leaky-)ReLU site stores the branch decisions it took (all rows), so that a
        # second oracle of another precision can be run with exactly these decisions (tests: fp32 vs fp64 truth).
        self.record = None
        # Optional dict: when set, named intermediate tensors of the forward are kept (with retain_grad) so that a test
        # can compare hidden activations and their gradients, not only the leaves.
        self.keep = None

    def _keep(self, name, t):
        if self.keep is not None:
            if t.requires_grad:
                t.retain_grad()
            self.keep[name] = t
        return t

    # ---- ops ---------------------------------------------------------------------------------------
    def chebyshev5(self, x, Lt, W, K):
        """models.py:69-103, including its layout shuffles."""
        N, M, Fin = x.shape
        x0 = x.permute(1, 2, 0).reshape(M, Fin * N)           # :81-82
        xs = [x0]
        if K > 1:
            x1 = torch.sparse.mm(Lt, x0)                        # :91
            xs.append(x1)
        for _ in range(2, K):
            x2 = 2 * torch.sparse.mm(Lt, x1) - x0               # :94
            xs.append(x2)
            x0, x1 = x1, x2
        xk = torch.stack(xs, 0).reshape(K, M, Fin, N)           # :97
        xk = xk.permute(3, 1, 2, 0).reshape(N * M, Fin * K)     # :98-99
        return (xk @ W).reshape(N, M, -1)                       # :102-103

    def _act(self, v, slope, site):
        """(leaky-)ReLU of the pre-activation v; honours an imposed branch mask for `site` if one was given."""
        pos = v > 0
        if self.masks is not None and site in self.masks:
            m = self.masks[site]
            rows = self.mask_rows.get(site)
            if rows is None:
                pos = m
            else:
                pos = pos.clone()
                pos[:, rows] = m
        if self.record is not None and site is not None:
            self.record[site] = pos.clone()
        return torch.where(pos, v, slope * v)

    def b1leakyrelu(self, x, b, site=None):
        """models.py:105-109 (tf.nn.leaky_relu default alpha=0.2)."""
        return self._act(x + b.reshape(1, 1, -1), 0.2, site)

    def b1tanh(self, x, b):
        """models.py:111-115."""
        return torch.tanh(x + b.reshape(1, 1, -1))

    def b1relu(self, x, b, site=None):
        """models.py:117-121."""
        return self._act(x + b.reshape(1, 1, -1), 0.0, site)

    def b2relu(self, x, b, site=None):
        """models.py:123-127: one bias per vertex and filter, b [1, M, F]."""
        return self._act(x + b.reshape(1, x.shape[1], x.shape[2]), 0.0, site)

    def cnp(self, x, Lt, Dm, W, b, K, brelu=None):
        """models.py:154-171: filter -> brelu -> pool."""
        brelu = brelu or self.b1leakyrelu
        return self.poolwT(brelu(self.chebyshev5(x, Lt, W, K), b), Dm)

    def udn(self, x, Lt, Um, W, b, K, brelu=None):
        """models.py:173-191: unpool -> filter (Laplacian of the finer level) -> brelu."""
        brelu = brelu or self.b1leakyrelu
        return brelu(self.chebyshev5(self.poolwT(x, Um), Lt, W, K), b)

    def poolwT(self, x, S):
        """models.py:129-152."""
        N, M, Fin = x.shape
        Mp = S.shape[0]
        xt = x.permute(1, 2, 0).reshape(M, Fin * N)
        xt = torch.sparse.mm(S, xt)
        return xt.reshape(Mp, Fin, N).permute(2, 0, 1)

    def fit_cond_dim(self, x, y):
        """models.py:813-832."""
        return y.reshape(x.shape[0], 1, -1) * torch.ones(x.shape[0], x.shape[1], y.shape[-1], dtype=x.dtype)

    def dense(self, x, P, scope, act=None, site=None):
        y = x @ P[scope + "/dense/kernel"] + P[scope + "/dense/bias"]
        if act == "leaky":
            y = self._act(y, 0.2, site)
        return y

    def gn(self, x, gamma, beta, G=32, eps=1e-5):
        """models.py:693-709."""
        xt = x.permute(0, 2, 1)
        N, C, V = xt.shape
        G = min(G, C)
        xg = xt.reshape(N, G, C // G, V)
        mean = xg.mean(dim=(2, 3), keepdim=True)
        var = ((xg - mean) ** 2).mean(dim=(2, 3), keepdim=True)
        xg = (xg - mean) / torch.sqrt(var + eps)
        out = xg.reshape(N, C, V) * gamma.reshape(1, C, 1) + beta.reshape(1, C, 1)
        return out.permute(0, 2, 1)

    # ---- network -----------------------------------------------------------------------------------
    def condition(self, y, P, name, nz_cond, nlayers, tag=""):
        """models.py:479-511."""
        scope = "condition_%s" % name
        if nlayers == 1:
            return self.dense(y, P, scope + "/fc1")
        y = self.dense(y, P, scope + "/fc1", act="leaky", site="cond_%s%s" % (name, tag))
        return self.dense(y, P, scope + "/fc2")

    def cond_embeddings(self, cond, cond2, P, tag=""):
        """models.py:284-286: pose net has nlayers=2 hard-coded, clothing net n_layer_cond."""
        y = self.condition(cond, P, "pose", self.cfg["nz_cond"], 2, tag)
        y2 = self.condition(cond2, P, "clo_label", self.cfg["nz_cond2"], self.cfg.get("n_layer_cond", 1), tag)
        return y, y2

    def encoder(self, x, P):
        """models.py:514-561 with use_res_block=0, cond_encoder=0 (all shipped configs)."""
        s = "generator/encoder/"
        for i in range(len(self.F)):
            sc = s + "encoder_conv%d" % (i + 1)
            x = self.chebyshev5(x, self.Lt[i], P[sc + "/weights"], self.K[i])     # cnp :164
            x = self.b1leakyrelu(x, P[sc + "/bias"], site="enc%d" % (i + 1))      # :166
            x = self._keep("enc_act%d" % (i + 1), self.poolwT(x, self.Dm[i]))     # :168
        if self.reduce_dim > 0:
            x = self._keep("enc_red", self.chebyshev5(x, self.Lt[-1], P[s + "1x1-conv/weights"], 1))   # :551
        x = x.reshape(x.shape[0], -1)                                             # :554
        z_mean = self.dense(x, P, s + "fc_mean")
        z_logvar = self.dense(x, P, s + "fc_var")
        return z_mean, z_logvar

    def res_block_affine(self, x, i, P, scope):
        """models.py:776-793."""
        x = self.poolwT(x, self.Um[-i - 1])
        Lt = self.Lt[-i - 2]
        x_gc = self.chebyshev5(x, Lt, P[scope + "/graph_conv/weights"], self.K[-i - 1])
        x_gc = self._act(x_gc, 0.0, "dec%d" % (i + 1))                             # tf.nn.relu, :785
        x_aff = self.chebyshev5(x, Lt, P[scope + "/affine/weights"], 1)
        return x_aff + x_gc

    def res_block_decoder(self, x_in, i, P, scope):
        """models.py:744-774."""
        x_unpooled = self.poolwT(x_in, self.Um[-i - 1])
        Lt = self.Lt[-i - 2]
        Fo = self.F[-i - 1]
        x = self._act(self.gn(x_unpooled, P[scope + "/group_norm/gamma"], P[scope + "/group_norm/beta"]), 0.0,
                      "gn%d_0" % (i + 1))                                          # tf.nn.relu, :752
        x = self.chebyshev5(x, Lt, P[scope + "/graph_linear_1/weights"], 1)
        x = self._act(self.gn(x, P[scope + "/group_norm_1/gamma"], P[scope + "/group_norm_1/beta"]), 0.0,
                      "gn%d_1" % (i + 1))
        x = self.chebyshev5(x, Lt, P[scope + "/graph_conv/weights"], self.K[-i - 1])
        x = self._act(self.gn(x, P[scope + "/group_norm_2/gamma"], P[scope + "/group_norm_2/beta"]), 0.0,
                      "gn%d_2" % (i + 1))
        x = self.chebyshev5(x, Lt, P[scope + "/graph_linear_2/weights"], 1)
        if x_unpooled.shape[-1] != Fo:
            x_unpooled = self.chebyshev5(x_unpooled, Lt, P[scope + "/graph_linear_input/weights"], 1)
        return x + x_unpooled

    def decoder_cond_vert(self, z_total, y, y2, P):
        """models.py:564-617 with use_res_block_dec=1."""
        s = "generator/decoder/"
        x = self._keep("dec_fc", self.dense(z_total, P, s + "fc1", act="leaky", site="dec_fc1"))   # :582
        x = x.reshape(x.shape[0], self.p[-1], -1)                                # :584
        if self.reduce_dim > 0:
            x = self.chebyshev5(x, self.Lt[-1], P[s + "1x1-conv/weights"], 1)     # :588
        x = torch.cat([x, self.fit_cond_dim(x, y), self.fit_cond_dim(x, y2)], -1)  # :591-594
        for i in range(len(self.F)):
            if self.cfg["affine"]:
                x = self.res_block_affine(x, i, P, s + "decoder_resblock_affine%d" % (i + 1))
            else:
                x = self.res_block_decoder(x, i, P, s + "decoder_resblock_cmr%d" % (i + 1))
            self._keep("dec_act%d" % (i + 1), x)
            x = torch.cat([x, self.fit_cond_dim(x, y), self.fit_cond_dim(x, y2)], -1)  # :606-609
        x = self.chebyshev5(x, self.Lt[0], P[s + "outputs/weights"], self.K[0])   # :612
        return x + P[s + "outputs/bias"]                                          # :615-616

    def generator(self, x, y, y2, eps, P):
        """models.py:620-645; eps is vae_sampling's random_normal made explicit (:194)."""
        z_mean, z_logvar = self.encoder(x, P)
        z = z_mean + torch.sqrt(torch.exp(z_logvar)) * eps                        # :195
        z_total = self._keep("z_total", torch.cat([z, y, y2], 1))                 # :641
        return self.decoder_cond_vert(z_total, y, y2, P), z_mean, z_logvar

    def discriminator(self, x, y, y2, P, tag=""):
        """models.py:648-678 (pred_map uses self.poly_order[-1], not Kd: :676)."""
        x = torch.cat([x, self.fit_cond_dim(x, y), self.fit_cond_dim(x, y2)], -1)
        for i in range(len(self.Dm_d)):
            sc = "discriminator/shared/conv%d" % (i + 1)
            x = self.chebyshev5(x, self.Lt_d[i], P[sc + "/weights"], self.Kd)     # cnp_d :803
            x = self.b1leakyrelu(x, P[sc + "/bias"], site="disc%d%s" % (i + 1, tag))
            x = self.poolwT(x, self.Dm_d[i])
        return self.chebyshev5(x, self.Lt_d[-1], P["discriminator/prediction_map/weights"], self.K[-1])

    # ---- losses (models.py:354-416, losses.py:9-25) ---------------------------------------------------
    @staticmethod
    def bce_logits(l, t):
        return (torch.clamp(l, min=0) - l * t + torch.log1p(torch.exp(-torch.abs(l)))).mean()

    def losses(self, x_hat, gt, z_mean, z_logvar, d_real, d_fake, P, edges, smooth=0.1):
        cfg = self.cfg
        d = x_hat - gt
        if self.masks is not None and "l1_sign" in self.masks:
            # |d| with the sign decisions of the implementation under test (same reason as the ReLU decisions: at
            # batch 64 a handful of the 1.3 M residuals lie within fp32 rounding of zero)
            # (three-valued: a residual that is exactly zero in the implementation under test has gradient 0 there)
            recon = (self.masks["l1_sign"].to(d.dtype) * d).mean()
        else:
            recon = d.abs().mean()                                                 # :358-360
        if self.record is not None:
            self.record["l1_sign"] = torch.sign(d.detach())
        latent = (-0.5 * (1 + z_logvar - z_mean ** 2 - torch.exp(z_logvar)).sum(1)).mean()  # :371-372
        e0 = torch.as_tensor(edges[:, 0].astype(np.int64))
        e1 = torch.as_tensor(edges[:, 1].astype(np.int64))
        ev = lambda a: a[:, e0] - a[:, e1]
        edge = torch.linalg.norm(ev(x_hat) - ev(gt), dim=-1).mean()                # losses.py:21-25
        reg = cfg["regularization"]
        # models.py:378: regularization * sum(l2_regularizer(regularization)(kernel)) over 'generator' dense kernels
        reg_g = 0.0
        for name in ("generator/encoder/fc_mean", "generator/encoder/fc_var", "generator/decoder/fc1"):
            reg_g = reg_g + reg * (P[name + "/dense/kernel"] ** 2).sum() / 2
        reg_g = reg * reg_g
        out = OrderedDict(recon=recon, latent=latent, edge=edge, reg_g=reg_g)
        if d_fake is not None:
            out["gan_g"] = self.bce_logits(d_fake, 1 - smooth)                     # :387
            out["loss_g"] = (out["gan_g"] * cfg["lambda_gan"] + recon * cfg["lambda_recon"] + edge * cfg["lambda_edge"]
                             + latent * cfg["lambda_latent"] + reg_g)              # :393-395
            if d_real is not None:
                out["gan_d"] = self.bce_logits(d_real, 1 - smooth) + self.bce_logits(d_fake, smooth)  # :388-390
                out["loss_d"] = out["gan_d"] * cfg["lambda_gan"]                   # :397 (reg_d = 0: no dense in D)
        return out


# --------------------------------------------------------------------------------------------------
# one training update (models.py:419-474)
# --------------------------------------------------------------------------------------------------
def lr_schedule(cfg, step):
    """models.py:426-442.  `step` = global_step value read when the update runs."""
    lr_g = cfg["lr"]
    lr_d = cfg["lr"] * cfg["lr_scaler"]
    ds = int(cfg["decay_steps"])
    if cfg.get("lr_warmup", False):
        warm = int(cfg["decay_steps"] * 8)
        if step < warm:
            return lr_g * step / warm, lr_d * step / warm
        k = math.floor((step - warm) / ds)
    else:
        k = math.floor(step / ds)
    return lr_g * cfg["decay_rate"] ** k, lr_d * cfg["decay_rate"] ** k


def g_var_names(P, optim_condnet=True):
    return [k for k in P if k.startswith("generator") or (optim_condnet and "condition" in k)]


def d_var_names(P):
    return [k for k in P if k.startswith("discriminator")]


ADAM_B1, ADAM_B2 = float(np.float32(0.9)), float(np.float32(0.999))    # TF casts the hyper-parameters to the variable's dtype


def adam_apply(p, m, v, g, lr, t, b1=ADAM_B1, b2=ADAM_B2, eps=1e-8):
    """One application of tf.train.AdamOptimizer (TF-1.13 defaults; the reference's `optimizer: adam` branch,
    models.py:450-451) to one variable: returns (p', m', v').  t = 1 for the first application.  TF folds the bias
    correction into the step size and adds eps to the UNcorrected sqrt(v) ("epsilon hat" in the Adam paper); beta1 /
    beta2 enter as float32 tensors, so (1 - beta2) is 1 - fl32(0.999) = 9.9998713e-4."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    return p - lr_t * m / (torch.sqrt(v) + eps), m, v


def train_update(oracle, P, mom, batch, step, edges, ref_compat=False, clip=5.0):
    """One optimiser application = what a single sess.run(op_train_*) does (models.py:460-472).

    batch: dict with x_g, gt, cond_g, cond2_g, eps, x_d, cond_d, cond2_d (torch tensors).
    ref_compat=True reproduces models.py:466 (the discriminator 'gradients' are its clipped variables);
    False applies the real discriminator gradients (clipped by their own global norm).
    Updates P and mom in place (plain tensors), returns the loss dict.
    """
    cfg = oracle.cfg
    Pg = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    y, y2 = oracle.cond_embeddings(batch["cond_g"], batch["cond2_g"], Pg, tag="_g")
    yd, y2d = oracle.cond_embeddings(batch["cond_d"], batch["cond2_d"], Pg, tag="_d")
    x_hat, zm, zl = oracle.generator(batch["x_g"], y, y2, batch["eps"], Pg)
    d_real = oracle.discriminator(batch["x_d"], yd, y2d, Pg, tag="_real")
    d_fake = oracle.discriminator(x_hat, y, y2, Pg, tag="_fake")
    L = oracle.losses(x_hat, batch["gt"], zm, zl, d_real, d_fake, Pg, edges)
    gn = g_var_names(Pg, cfg.get("optim_condnet", True))
    dn = d_var_names(Pg)
    grads_g = torch.autograd.grad(L["loss_g"], [Pg[k] for k in gn], retain_graph=True, allow_unused=True)
    grads_g = [g if g is not None else torch.zeros_like(Pg[k]) for g, k in zip(grads_g, gn)]
    if ref_compat:
        grads_d = [Pg[k].detach() for k in dn]                                     # models.py:466
    else:
        grads_d = torch.autograd.grad(L["loss_d"], [Pg[k] for k in dn], allow_unused=True)
        grads_d = [g if g is not None else torch.zeros_like(Pg[k]) for g, k in zip(grads_d, dn)]
    lr_g, lr_d = lr_schedule(cfg, step)
    adam = cfg.get("optimizer", "sgd") == "adam"
    if adam:
        # tf.train.AdamOptimizer (models.py:450-451; TF-1.13 defaults): `mom` carries the first moments under the variable
        # names, the second moments under "adam_v/<name>" and the number of applications so far under "adam_t"
        t = int(mom.get("adam_t", torch.zeros(())).item()) + 1
        mom["adam_t"] = torch.tensor(float(t))
    for names, grads, lr in ((gn, grads_g, lr_g), (dn, grads_d, lr_d)):
        norm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).item()
        coef = clip / max(norm, clip)
        for k, g in zip(names, grads):
            if adam:
                gc = coef * g.detach()
                v = mom.get("adam_v/" + k)
                v = torch.zeros_like(gc) if v is None else v
                P[k], mom[k], mom["adam_v/" + k] = adam_apply(P[k], mom[k], v, gc, lr, t)
            else:
                mom[k] = cfg["momentum"] * mom[k] + coef * g.detach()
                P[k] = P[k] - lr * mom[k]
    out = {k: float(v) for k, v in L.items()}
    out["grads"] = {k: g.detach() for k, g in zip(gn + dn, list(grads_g) + list(grads_d))}
    out["mom"] = {k: mom[k].detach() for k in gn + dn}
    if adam:
        out["adam_v"] = {k: mom["adam_v/" + k].detach() for k in gn + dn}
    out["x_hat"] = x_hat.detach()
    return out
