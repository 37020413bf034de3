"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle (tests/parity.py).
Tolerance: 1e-4 relative fp32 (BASELINE.json north_star)."""
import pytest
import torch

import parity

pytestmark = pytest.mark.gpu


def _assert_all(res, tol=parity.TOL):
    bad = {k: v for k, v in res.items() if not v < tol}
    assert not bad, "parity failures (rel err): %s" % bad


@pytest.fixture(scope="module")
def cfg():
    from cape_b200.params import NZ64_AFFINE
    return dict(NZ64_AFFINE, decay_steps=10)


def test_golden_vectors(hierarchy):
    """BASELINE configs[0] (single Chebyshev K=6 layer on the 6890x3 template, batch 1) + fused cnp + unpool
    against the committed golden outputs."""
    _assert_all(parity.golden_ops(hierarchy))


def test_dense_layers():
    _assert_all(parity.gemm_cases())


def test_chebyshev_forward_and_gradients(hierarchy):
    """Every shape class of the path: pooled encoder conv, unpooled decoder conv, K=3 discriminator conv, 1x1,
    thin (Fout=3, Fout=1) and first (Fin=3) layers -- forward, dx, dW, db."""
    _assert_all(parity.cheb_grad_cases(hierarchy))


def test_plain_operand_kernel(hierarchy):
    """1x1 convs / plain-tensor terms on the TMA-fed kernel: odd widths, > 512 columns, multi-tile, all gradients."""
    _assert_all(parity.plain_operand_cases(hierarchy))


def test_plain_operand_kernel_without_presplit_weights(hierarchy):
    """Same calls with experiment knob 15: the weight lo tiles derived on chip by the converter warps (the path a caller
    takes who passes no cape_term.wT_lo) instead of fetched by TMA from the pre-split copy; plus the precise mode."""
    from cape_b200 import _lib
    lib = _lib.load()
    prev = lib.cape_set_tuning(15, 1)
    try:
        _assert_all(parity.plain_operand_cases(hierarchy))
        res = parity.precise_vs_truth(hierarchy)
        assert res["precise L8 1024->512 (max-rel vs fp64)"] < 4e-6, res
    finally:
        lib.cape_set_tuning(15, prev)


def test_precise_accumulation(hierarchy):
    """cape_conv_args.precise: split tensor-core accumulation chains -- close to fp32 SIMT accuracy, and at least three
    times closer to the float64 truth than the single-chain default on a 1024-long reduction."""
    res = parity.precise_vs_truth(hierarchy)
    assert res["precise L8 1024->512 (max-rel vs fp64)"] < 4e-6, res
    assert res["precise L8 1024->512 (max-rel vs fp64)"] * 3 < res["default L8 1024->512 (max-rel vs fp64)"], res
    assert res["precise L8 512->64 (max-rel vs fp64)"] < 3e-6, res


def test_apply_operators(hierarchy):
    """cape_apply against scipy sparse products (float64)."""
    _assert_all(parity.apply_cases(hierarchy), tol=2e-6)


def test_group_norm():
    _assert_all(parity.gn_case())
    _assert_all(parity.gn_case(N=3, rows=6890, C=32, seed=1))


def test_train_step_matches_oracle(hierarchy, cfg):
    """Full VAE+GAN update (enc+dec+disc fwd/bwd, losses, clip, momentum): x_hat, 5 loss terms, every gradient and
    every post-update parameter."""
    res = parity.train_step(hierarchy, cfg, N=2)
    assert res["x_hat (vertex-L2)"] < 1e-4
    _assert_all(res)


@pytest.mark.xfail(strict=False, reason="written after the round's GPU budget was spent: not yet confirmed on a GPU (an "
                   "earlier version that also compared gradient-like parameters failed, see tests/parity.py); an XPASS "
                   "in the report is the confirmation, a failure must not stop `pytest -x`")
def test_train_step_matches_the_reference_golden_file(hierarchy):
    """The CUDA path against numbers produced by the REFERENCE's own lib/models.py (executed on the TF-API shim,
    tests/golden/make_ref_golden.py): x_hat, the five loss terms and the discriminator's post-update parameters (the
    lib/models.py:466 update) of one full update.  No oracle in between."""
    _assert_all(parity.reference_golden_update(hierarchy))


def test_train_step_reference_initialisers_vs_float64_truth(hierarchy, cfg):
    """The reference's own initialisers (no calibration: glorot fc_mean/fc_var on N(0,1) inputs, logvar up to +-10,
    KL term ~1e4).  Truth = the float64 oracle; the CUDA path must be within 1e-4 of it, or at least as close as twice
    what a plain fp32 CPU implementation (the fp32 oracle) achieves on the same update."""
    res = parity.train_step(hierarchy, cfg, N=2, fc_scale=1.0, truth=True)
    bad = {k: v for k, v in res.items() if not v[0] < max(parity.TOL, 2.0 * v[1])}
    assert not bad, "further from the fp64 truth than an fp32 CPU implementation (err, fp32-oracle err): %s" % bad


def test_train_step_unmasked_forward(hierarchy, cfg, capsys):
    """Forward-side quantities (x_hat, the five loss terms) against an oracle that takes its OWN branch decisions, so
    a wrong sign/branch in an epilogue cannot hide behind the imposed masks; the gradient errors of that unmasked
    comparison are printed (they contain the handful of legitimately flipped near-zero units), not asserted."""
    res = parity.train_step(hierarchy, cfg, N=2, fc_scale=1.0, report_unmasked=True)
    fwd = {k: v for k, v in res.items() if k.startswith("unmasked fwd")}
    assert len(fwd) == 7
    _assert_all(fwd)
    with capsys.disabled():
        print("\nunmasked gradient errors (informational):")
        for k, v in res.items():
            if k.startswith("unmasked grad"):
                print("  %-75s %.2e" % (k[len("unmasked grad "):], v))
    _assert_all({k: v for k, v in res.items() if not k.startswith("unmasked")})


def test_three_consecutive_updates(hierarchy, cfg):
    """Momentum != 0, warm-up learning rates, refreshed K-major / tf32-low weight copies: three updates in a row with
    fresh batches, every update compared with the oracle carrying its own state (lib/models.py:460-472)."""
    _assert_all(parity.train_step(hierarchy, cfg, N=2, nsteps=3, fc_scale=1.0))


def test_three_consecutive_updates_graphs(hierarchy, cfg):
    """Same through the two captured CUDA graphs (the learning rate and inputs change under the graphs)."""
    _assert_all(parity.train_step(hierarchy, cfg, N=3, nsteps=3, use_graph=True, seed=11))


def test_train_step_full_batch_c3(hierarchy, cfg):
    """BASELINE configs[2] at its own size: batch 64 (more 128-row tiles than SMs, persistent loops, split-K shapes of
    the benchmark, side-stream weight gradients, CUDA graphs) against the oracle."""
    _assert_all(parity.train_step(hierarchy, cfg, N=64, use_graph=True, fc_scale=1.0))


def test_generator_forward_c2(hierarchy, cfg):
    """BASELINE configs[1]: encoder+decoder forward at batch 32 against the oracle."""
    _assert_all(parity.generator_forward(hierarchy, cfg, N=32))


def test_train_step_reference_quirks(hierarchy, cfg):
    """ref_compat=True reproduces lib/models.py:466 (discriminator 'gradients' = its clipped variables)."""
    _assert_all(parity.train_step(hierarchy, cfg, N=2, ref_compat=True))


def test_adam_kernel():
    """cape_adam_clip_update == tf.train.AdamOptimizer's update rule (float64 formula), clip active and inactive."""
    _assert_all(parity.adam_kernel_case(), tol=1e-5)


def test_train_step_adam(hierarchy, cfg):
    """`optimizer: adam` (lib/models.py:449-451): one update -- gradients, both moment slots and the parameters where
    the gradient is resolved (see parity.train_step) against the oracle's Adam."""
    _assert_all(parity.train_step(hierarchy, dict(cfg, optimizer="adam"), N=2))


def test_tensor_core_path_matches_simt(hierarchy):
    """The tcgen05 3xTF32 contraction and the fp32 FFMA contraction are two implementations of one entry point."""
    _assert_all(parity.tc_vs_simt(hierarchy), tol=2e-5)


def test_train_step_simt_only(hierarchy, cfg):
    """Same full-step parity with the tensor-core path switched off (every layer on the fp32 SIMT kernels)."""
    prev = parity.set_tensor_cores(False)
    try:
        _assert_all(parity.train_step(hierarchy, cfg, N=2))
    finally:
        parity.set_tensor_cores(prev)


def test_train_step_patch_vertex_order(hierarchy, cfg):
    """reorder=True: hidden activations kept in patch order (topology.patch_order) -- a layout change only."""
    _assert_all(parity.train_step(hierarchy, cfg, N=2, reorder=True))


def test_train_step_regathered_weight_gradient(hierarchy, cfg, monkeypatch):
    """CAPE_DW_STASH=0: cape_cheb_dw gathers the basis again instead of contracting the stashed copies."""
    monkeypatch.setenv("CAPE_DW_STASH", "0")
    _assert_all(parity.train_step(hierarchy, cfg, N=2))


def test_train_step_odd_batch(hierarchy, cfg):
    """Batch that does not divide the 128-row tiles; other seed."""
    _assert_all(parity.train_step(hierarchy, cfg, N=5, seed=7))


def test_train_step_groupnorm_decoder(hierarchy):
    """BASELINE configs[4]: CAPE nz18_pose24_clotype8, non-affine decoder (GroupNorm residual blocks,
    lib/models.py:744-774) -- the plain chebyshev5 path; full update vs the oracle."""
    from cape_b200.params import NZ18_PLAIN
    _assert_all(parity.train_step(hierarchy, dict(NZ18_PLAIN, decay_steps=10), N=2))


def test_train_step_groupnorm_decoder_many_tiles(hierarchy):
    """The nz18 / GroupNorm model at batch 24: more 128-row tiles than SMs in every layer (persistent loops and
    column groups of the plain-operand kernel on the 544 / 288 / 160-wide linear layers), graph-replayed."""
    from cape_b200.params import NZ18_PLAIN
    _assert_all(parity.train_step(hierarchy, dict(NZ18_PLAIN, decay_steps=10), N=24, use_graph=True))


def test_size_independent_properties(hierarchy, cfg):
    """Full-size (batch 64) checks that need no oracle: linearity of the conv in x and W, batch-permutation
    equivariance of the generator, CUDA-graph replay == eager."""
    import numpy as np
    from cape_b200 import ops
    from cape_b200.network import CapeNetwork
    from cape_b200.synthetic import make_batch
    h = hierarchy
    g = torch.Generator(device="cuda").manual_seed(0)
    x1 = torch.randn(64, 6890, 64, device="cuda", generator=g)
    x2 = torch.randn(64, 6890, 64, device="cuda", generator=g)
    W = torch.randn(128, 64, device="cuda", generator=g) * 0.1
    f = lambda x, w: ops.chebyshev5(x, h["L"][1], w, 2, pool=h["D"][1])
    lhs = f(2.0 * x1 - 3.0 * x2, W)
    rhs = 2.0 * f(x1, W) - 3.0 * f(x2, W)
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) < 1e-5
    N = 64
    net = CapeNetwork(h["L"], h["D"], h["U"], h["L_d"], h["D_d"], cfg, N)
    b = {k: torch.from_numpy(v) for k, v in make_batch(N, cfg["nz"], seed=3).items()}
    net.set_inputs(b["x_g"], b["cond_g"], b["cond2_g"], b["eps"], b["x_d"], b["cond_d"], b["cond2_d"])
    y = net.forward_generator().clone()
    perm = torch.from_numpy(np.random.RandomState(0).permutation(N))
    net.set_inputs(b["x_g"][perm], b["cond_g"][perm], b["cond2_g"][perm], b["eps"][perm])
    yp = net.forward_generator().clone()
    assert float((yp - y[perm.cuda()]).abs().max() / y.abs().max()) < 1e-5
    # eager step == graph-replayed step (same inputs, update disabled)
    net.set_inputs(b["x_g"], b["cond_g"], b["cond2_g"], b["eps"], b["x_d"], b["cond_d"], b["cond2_d"])
    net.train_step(step=100, update=False)
    g_eager = net.PG.grad.clone()
    net.capture_graphs()
    net.train_step(step=100, update=False, use_graph=True)
    torch.cuda.synchronize()
    d = float((net.PG.grad - g_eager).abs().max() / g_eager.abs().max())
    assert d < 1e-5, d


def test_global_norm_is_deterministic():
    """cape_sumsq reduces in a fixed order: the clip factor of the update must be bit-identical on every data-parallel
    replica (and in every run), otherwise replicas that clip drift apart by an ulp per step."""
    import ctypes as C
    from cape_b200 import _lib
    lib = _lib.load()
    g = torch.randn(16_285_668, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)) * 0.01
    out = torch.zeros(8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for i in range(8):
        _lib.check(lib.cape_sumsq(C.c_void_p(g.data_ptr()), g.numel(), C.c_void_p(out[i:].data_ptr()), st))
    torch.cuda.synchronize()
    v = out.cpu().numpy()
    assert (v == v[0]).all(), v
    ref = float((g.double() ** 2).sum())
    assert abs(float(v[0]) - ref) < 1e-5 * ref
