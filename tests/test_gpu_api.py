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


def test_demo_simple_from_a_tensorflow_checkpoint(hierarchy, tmp_path):
    """The user-visible product path (run_simple_demo.py + demos.py:339-406): weights come from a TensorFlow-format
    checkpoint (the reference's Saver format, here written by cape_b200.tf_checkpoint from synthetic weights), the
    model restores it by itself, the demo decodes three latent samples for each of the four clothing types,
    de-normalises with trainset_stats, masks the non-clothing vertices, adds the template and writes OBJ files --
    checked against the oracle's decoder + the same post-processing in numpy."""
    from oracle import cape_oracle as O
    from cape_b200 import tf_checkpoint, topology as T
    from cape_b200.demos import demo_simple, filter_cloth_pose, read_obj
    h = hierarchy
    m = _model(h, batch_size=4)
    params = parity.calibrated_params(m.net.specs, 21)
    ck = tmp_path / "ckpt" / m.name
    tf_checkpoint.write_checkpoint(str(ck / "model.ckpt-777"), dict(params, global_step=np.asarray(777, np.int64)))
    m.checkpoint_dir = str(tmp_path / "ckpt")
    demo = demo_simple(m, m.name, random_seed=5, results_dir=str(tmp_path / "out"))
    got = demo.sample_vary_clotype()
    assert m._weights_source == "checkpoint" and m.global_step == 777
    # the same computation with the oracle
    o = O.Oracle(h["L"], h["D"], h["U"], h["L_d"], h["D_d"], m.cfg)
    P = {k: torch.from_numpy(v) for k, v in params.items()}
    t = torch.from_numpy
    rot = filter_cloth_pose(T.demo_pose_params()[0])[0].astype(np.float32)
    clo = np.eye(4, dtype=np.float32)
    y, y2 = o.cond_embeddings(t(np.repeat(rot[None], 4, 0)), t(clo), P)
    np.random.seed(5)
    z = np.random.normal(size=(3, 64))
    mean, std = T.trainset_stats()
    tv, tf_ = T.template_mesh()
    keep = T.clothing_verts_idx()
    for i, name in enumerate(["shortlong", "shortshort", "longshort", "longlong"]):
        zt = np.concatenate([z, np.repeat(y[:1].numpy(), 3, 0), np.repeat(y2[i:i + 1].numpy(), 3, 0)], 1).astype(np.float32)
        pred = o.decoder_cond_vert(t(zt), y[:1].repeat(3, 1), y2[i:i + 1].repeat(3, 1), P).numpy()
        pred = pred * std + mean
        want = np.zeros_like(pred)
        want[:, keep] = pred[:, keep]
        want = want + tv
        assert parity.vertex_l2(got[name] - tv, want - tv) < 1e-4
        v, f = read_obj(str(tmp_path / "out" / ("%s_0002.obj" % name)))
        assert np.array_equal(f, tf_) and np.abs(v - got[name][2]).max() < 1e-6
    assert len(list((tmp_path / "out").iterdir())) == 12


def test_tensorflow_checkpoint_export_roundtrip(hierarchy, tmp_path):
    """save_tf -> restore: weights, Momentum slots and global_step travel through the TensorFlow bundle format."""
    from cape_b200 import tf_checkpoint
    m = _model(hierarchy, batch_size=2)
    m.checkpoint_dir = str(tmp_path)
    m.load_weights(parity.calibrated_params(m.net.specs, 9))
    m.net.PG.mom.normal_()
    m.global_step = 31
    before, mom = m.net.get_params(), m.net.PG.export(m.net.PG.mom)
    prefix = m.save_tf(5)
    names = {n for n, _, _ in tf_checkpoint.list_variables(prefix)}
    assert "generator/decoder/outputs/weights/Momentum" in names and "global_step" in names
    m2 = _model(hierarchy, batch_size=2)
    m2.checkpoint_dir = str(tmp_path)
    assert m2.restore() == prefix and m2.global_step == 31
    after = m2.net.get_params()
    assert all(np.array_equal(before[k], after[k]) for k in before)
    mom2 = m2.net.PG.export(m2.net.PG.mom)
    assert all(np.array_equal(mom[k], mom2[k]) for k in mom)


class _Data:
    pass


def _synthetic_data(n_train, n_val, seed=0):
    from cape_b200.synthetic import make_batch
    b = make_batch(n_train + n_val, 64, seed=seed)
    d = _Data()
    d.vertices_train, d.cond1_train, d.cond2_train = b["x_g"][:n_train], b["cond_g"][:n_train], b["cond2_g"][:n_train]
    d.vertices_val, d.cond1_val, d.cond2_val = b["x_g"][n_train:], b["cond_g"][n_train:], b["cond2_g"][n_train:]
    return d


@pytest.mark.parametrize("device_dataset", [True, False])
def test_fit_runs_the_reference_loop(hierarchy, tmp_path, device_dataset):
    """CAPE.fit with ref_compat=True on a synthetic data_wrapper: two loop steps = four G+D updates (both sess.run calls
    of a loop step apply both optimisers, lib/models.py:470-472,905-906), global_step advances by 2 per update -> 8;
    every update is replayed by the oracle on the batch the loop actually staged (batches assembled on the GPU from
    indices, or on the host) and the final weights and momentum compared; a checkpoint is written."""
    h = hierarchy
    N = 2
    m = _model(h, batch_size=N)
    m.restart, m.num_epochs, m.checkpoint_dir, m.device_dataset = True, 1, str(tmp_path), device_dataset
    assert m.ref_compat
    params = parity.calibrated_params(m.net.specs, 31)
    m.load_weights(params)
    data = _synthetic_data(4, 2, seed=3)
    net, rec = m.net, []
    orig = net.train_step

    def spy(step=None, **kw):
        batch = dict(x_g=net.in_x.cpu().clone(), cond_g=net.in_cond[N:].cpu().clone(), cond2_g=net.in_cond2[N:].cpu().clone(),
                     eps=net.in_eps.cpu().clone(), x_d=net.xcat[:N].cpu().clone(), cond_d=net.in_cond[:N].cpu().clone(),
                     cond2_d=net.in_cond2[:N].cpu().clone())
        batch["gt"] = batch["x_g"]
        r = orig(step=step, **kw)
        torch.cuda.synchronize()
        rec.append((step, batch) + parity.cuda_masks(net, h, N))
        return r

    net.train_step = spy
    np.random.seed(17)
    losses, _ = m.fit(data)
    net.train_step = orig
    assert m.global_step == 8 and [r[0] for r in rec] == [0, 2, 4, 6]
    assert len(losses) == 1 and np.isfinite(losses[0])
    assert any(f.startswith("model-") for f in __import__("os").listdir(str(tmp_path / m.name)))
    # the staged batches are rows of the training split (the index deques of the loop)
    rows = {data.vertices_train[i].tobytes() for i in range(4)}
    assert all(b["x_g"][j].numpy().tobytes() in rows and b["x_d"][j].numpy().tobytes() in rows
               for _, b, _, _ in rec for j in range(N))
    o_params = {k: np.asarray(v, np.float32) for k, v in params.items()}
    o_mom = {k: np.zeros_like(v) for k, v in o_params.items()}
    for step, batch, masks, mrows in rec:
        _, o_params, o_mom = parity._oracle_update(h, m.cfg, o_params, o_mom, batch, step, torch.float32, True, masks, mrows)
    got_p = net.get_params()
    got_m = {**net.PG.export(net.PG.mom), **net.PD.export(net.PD.mom)}
    bad = {k: parity.rel(got_p[k], v) for k, v in o_params.items() if not parity.rel(got_p[k], v) < 1e-4}
    bad.update({"mom " + k: parity.rel(got_m[k], v) for k, v in o_mom.items() if not parity.rel(got_m[k], v) < 1e-4})
    assert not bad, bad


def test_device_dataset_batches_equal_host_batches(hierarchy):
    """load_data.DeviceDataset.stage (gather kernel, indices only over PCIe) fills the input buffers exactly like
    set_inputs with numpy fancy-indexing; BodyData normalises like lib/load_data.py:103-127."""
    from cape_b200.load_data import BodyData, DeviceDataset
    from cape_b200.network import CapeNetwork
    from cape_b200.params import NZ64_AFFINE
    rng = np.random.RandomState(0)
    raw = rng.normal(1.0, 2.0, size=(12, 6890, 3))
    pose = rng.normal(size=(12, 24, 9))                                   # full poses: reduced to the 14 clothing joints
    clo = np.eye(4)[rng.randint(0, 4, 12)]
    bd = BodyData(2, raw[:10], pose[:10], raw[10:], pose[10:], None, clo[:10], clo[10:])
    assert bd.vertices_train.shape == (8, 6890, 3) and bd.vertices_val.shape == (2, 6890, 3)
    assert bd.cond1_train.shape == (8, 126) and bd.cond1_train.dtype == np.float32
    assert np.allclose(bd.vertices_train.mean(0), 0, atol=1e-5) and np.allclose(bd.vertices_train.std(0), 1, atol=1e-4)
    assert np.allclose(bd.vertices_test, (raw[10:] - raw[:8].mean(0)) / raw[:8].std(0), atol=1e-5)
    N = 3
    net = CapeNetwork(hierarchy["L"], hierarchy["D"], hierarchy["U"], hierarchy["L_d"], hierarchy["D_d"],
                      dict(NZ64_AFFINE), N)
    ds = DeviceDataset.from_body_data(bd, net.device)
    idx_g, idx_d = [5, 0, 7], [2, 2, 6]
    eps = rng.normal(size=(N, 64)).astype(np.float32)
    ds.stage(net, idx_g, idx_d, eps)
    torch.cuda.synchronize()
    assert np.array_equal(net.in_x.cpu().numpy(), bd.vertices_train[idx_g])
    assert np.array_equal(net.xcat[:N].cpu().numpy(), bd.vertices_train[idx_d])
    assert np.array_equal(net.in_cond.cpu().numpy(), np.concatenate([bd.cond1_train[idx_d], bd.cond1_train[idx_g]]))
    assert np.array_equal(net.in_cond2.cpu().numpy(), np.concatenate([bd.cond2_train[idx_d], bd.cond2_train[idx_g]]))
    assert np.array_equal(net.in_eps.cpu().numpy(), eps)


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
    assert np.isfinite(a).all() and np.isfinite(b).all()
    # same weights, same batch in the first step: equal up to the order of the atomic accumulations (loss / column-sum /
    # condition-gradient kernels); the later steps have been through updates that amplify those last bits
    assert np.allclose(a[0], b[0], rtol=1e-5, atol=1e-7), (a[0], b[0])
    assert np.allclose(a, b, rtol=2e-3, atol=1e-5), (a, b)
    assert not np.array_equal(a[0], a[1])                     # the two batches differ


def test_train_step_on_a_generated_4_layer_hierarchy(hierarchy):
    """`--num_conv_layers 4` (main.py:31-32,56-57): the hierarchy is generated from the template mesh by
    cape_b200.mesh_sampling (the reference needs psbody for it and ships fixtures for 8 layers only), the model built on
    it takes a full VAE+GAN update, checked against the oracle built on the same generated operators."""
    from cape_b200 import main as M
    from cape_b200.params import NZ64_AFFINE
    L, D, U, p = M.build_hierarchy(num_conv_layers=4, ds_factor=2)
    assert p == [6890, 6890, 3445, 3445, 3445]
    h = dict(L=L, D=D, U=U, p=p, L_d=hierarchy["L_d"], D_d=hierarchy["D_d"])
    cfg = dict(NZ64_AFFINE, F=[64, 128, 128, 64], K=[2] * 4, decay_steps=10)
    res = parity.train_step(h, cfg, N=2)
    bad = {k: v for k, v in res.items() if not v < parity.TOL}
    assert not bad, bad
