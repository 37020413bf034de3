"""The reference's model-level API (lib/models.py:931-1174) on the GPU engine vs the oracle."""
import numpy as np
import pytest
import torch

import parity

pytestmark = pytest.mark.gpu


def _model(hierarchy, batch_size):
    from cape_b200.config_parser import model_params, parse_config
    from cape_b200.models import CAPE
    import tempfile, os
    yaml = ("nz: 64\nnz_cond: 32\nnz_cond2: 32\naffine: 1\nlr_warmup: 1\nname: api_test\nbatch_size: %d\n"
            "mode: demo\n" % batch_size)
    d = tempfile.mkdtemp()
    fn = os.path.join(d, "c.yaml")
    open(fn, "w").write(yaml)
    args, _ = parse_config(["--config", fn])
    p = model_params(args)
    p["p"] = hierarchy["p"]
    h = hierarchy
    m = CAPE(L=h["L"], D=h["D"], U=h["U"], L_d=h["L_d"], D_d=h["D_d"], **p)
    m.build_graph(m.input_num_verts, m.nn_input_channel, phase="demo")
    return m


def test_encode_decode_predict_match_oracle(hierarchy):
    from oracle import cape_oracle as O
    from cape_b200.synthetic import make_batch
    h = hierarchy
    m = _model(h, batch_size=4)
    specs = m.net.specs
    params = parity.calibrated_params(specs, 11)
    m.load_weights(params)
    n = 6                                                     # 1.5 batches: exercises the zero padding
    b = make_batch(n, 64, seed=5)
    o = O.Oracle(h["L"], h["D"], h["U"], h["L_d"], h["D_d"], m.cfg)
    P = {k: torch.from_numpy(v) for k, v in params.items()}
    t = torch.from_numpy
    y, y2 = o.cond_embeddings(t(b["cond_g"]), t(b["cond2_g"]), P)
    zm, zl = o.encoder(t(b["x_g"]), P)
    # encode / encode_only_condition
    gm, gl, gc, gc2 = m.encode(b["x_g"], b["cond_g"], b["cond2_g"])
    assert parity.rel(gm, zm.numpy()) < 1e-4 and parity.rel(gl, zl.numpy()) < 1e-4
    assert parity.rel(gc, y.numpy()) < 1e-5 and parity.rel(gc2, y2.numpy()) < 1e-5
    c1, c2 = m.encode_only_condition(b["cond_g"], b["cond2_g"])
    assert np.array_equal(c1, gc) and np.array_equal(c2, gc2)
    # decode from an explicit latent code
    z = np.random.RandomState(0).normal(size=(n, 64)).astype(np.float32)
    z_total = np.concatenate([z, gc, gc2], 1)
    want = o.decoder_cond_vert(t(z_total), y, y2, P).numpy()
    got = m.decode(z_total, gc, gc2)
    assert parity.vertex_l2(got, want) < 1e-4
    # one condition, many samples (demos.py usage, lib/models.py:1152-1153)
    got1 = m.decode(np.concatenate([z[:3], np.repeat(gc[:1], 3, 0), np.repeat(gc2[:1], 3, 0)], 1), gc[:1], gc2[:1])
    want1 = o.decoder_cond_vert(t(np.concatenate([z[:3], np.repeat(gc[:1], 3, 0), np.repeat(gc2[:1], 3, 0)], 1)),
                                y[:1].repeat(3, 1), y2[:1].repeat(3, 1), P).numpy()
    assert parity.vertex_l2(got1, want1) < 1e-4
    # predict draws eps from the model's RandomState(seed) per (padded) batch
    m.rng = np.random.RandomState(123)
    rng = np.random.RandomState(123)
    eps = np.concatenate([rng.normal(size=(4, 64)), rng.normal(size=(4, 64))]).astype(np.float32)[:8]
    preds, lr, ll, le = m.predict(b["x_g"], b["cond_g"], b["cond2_g"], labels=b["x_g"])
    xw = []
    for s in (slice(0, 4), slice(4, 6)):
        e = eps[s.start: s.start + (s.stop - s.start)]
        xh, _, _ = o.generator(t(b["x_g"][s]), y[s], y2[s], t(e), P)
        xw.append(xh.numpy())
    assert parity.vertex_l2(preds, np.concatenate(xw)) < 1e-4
    assert np.isfinite([lr, ll, le]).all()


def test_checkpoint_roundtrip(hierarchy, tmp_path):
    m = _model(hierarchy, batch_size=2)
    m.checkpoint_dir = str(tmp_path)
    before = m.net.get_params()
    m.global_step = 42
    fn = m.save(7)
    m.net.set_params({k: v * 0 for k, v in before.items()})
    m.restore()
    after = m.net.get_params()
    assert m.global_step == 42
    assert all(np.array_equal(before[k], after[k]) for k in before)
    assert set(np.load(fn).files) >= set(before)          # keyed by the reference's TF variable names
    assert m.get_var("generator/decoder/outputs/bias").shape == (1, 6890, 3)


def test_inference_entry_points_restore_the_checkpoint(hierarchy, tmp_path):
    """encode/decode on a freshly built model restore the newest checkpoint like the reference's _get_session
    (lib/models.py:209-215) -- save -> new CAPE -> decode reproduces the saved model's output -- and refuse to run on
    random initialisers when there is none."""
    m = _model(hierarchy, batch_size=2)
    m.checkpoint_dir = str(tmp_path)
    m.load_weights(parity.calibrated_params(m.net.specs, 5))
    z = np.random.RandomState(1).normal(size=(2, 128)).astype(np.float32)
    want = m.decode(z, z[:, 64:96], z[:, 96:])
    m.save(3)
    m2 = _model(hierarchy, batch_size=2)
    m2.checkpoint_dir = str(tmp_path)
    assert m2._weights_source == "init"
    got = m2.decode(z, z[:, 64:96], z[:, 96:])
    assert m2._weights_source == "checkpoint" and np.array_equal(got, want)
    m3 = _model(hierarchy, batch_size=2)
    m3.checkpoint_dir = str(tmp_path / "empty")
    with pytest.raises(FileNotFoundError):
        m3.decode(z, z[:, 64:96], z[:, 96:])
    m3.name = None                                  # config_parser's default: checkpoints directly under the folder
    assert m3._get_path("x") == "x/"


def test_prefetched_inputs_equal_direct_inputs(hierarchy):
    """prefetch_inputs/commit_inputs (copy stream + staging buffers, what bench.py's end-to-end loop uses) feed the
    step the same batch as set_inputs: two alternating batches, same losses (up to the last bits: the loss and
    column-sum kernels accumulate with atomics)."""
    from cape_b200.network import CapeNetwork
    from cape_b200.params import NZ64_AFFINE, param_specs
    from cape_b200.synthetic import make_batch
    h, cfg, N = hierarchy, dict(NZ64_AFFINE), 2
    specs = param_specs(cfg, [l.shape[0] for l in h["L"]], [l.shape[0] for l in h["L_d"]])
    params = parity.calibrated_params(specs, 3)
    order = ("x_g", "cond_g", "cond2_g", "eps", "x_d", "cond_d", "cond2_d")
    batches = [[torch.from_numpy(make_batch(N, cfg["nz"], seed=s)[k]).pin_memory() for k in order] for s in (1, 2)]

    def run(prefetch):
        net = CapeNetwork(h["L"], h["D"], h["U"], h["L_d"], h["D_d"], cfg, N, params=params)
        out = []
        if prefetch:
            net.prefetch_inputs(*batches[0])
        for i in range(4):
            if prefetch:
                net.commit_inputs()
                net.prefetch_inputs(*batches[(i + 1) % 2])
            else:
                net.set_inputs(*batches[i % 2])
            net.train_step(step=10 + 2 * i)
            out.append(net.losses.cpu().numpy().copy())
        return np.stack(out)

    a, b = run(False), run(True)
    assert np.isfinite(a).all() and np.allclose(a, b, rtol=1e-4, atol=1e-6)
    assert not np.array_equal(a[0], a[1])                     # the two batches differ
