"""Host-side pieces around the hot path that need no GPU: BodyData (lib/load_data.py:35-127), the demo helpers
(lib/utils.py:38-62, demos.py:394-402), OBJ io, the packed demo assets."""
import numpy as np
import pytest


def test_body_data_matches_the_reference_recipe(tmp_path):
    from cape_b200.load_data import BodyData
    rng = np.random.RandomState(1)
    verts = rng.normal(0.5, 3.0, size=(10, 50, 3))
    pose = rng.normal(size=(10, 24, 9))
    clo = np.eye(4)[rng.randint(0, 4, 10)]
    tv, tp, tc = rng.normal(size=(3, 50, 3)), rng.normal(size=(3, 216)), np.eye(4)[[0, 1, 2]]
    # files, as main.py passes them (load_data.py:57-60), and arrays must give the same object
    fn = {}
    for k, a in dict(v=verts, p=pose, c=clo, tv=tv, tp=tp, tc=tc).items():
        fn[k] = str(tmp_path / (k + ".npy"))
        np.save(fn[k], a)
    a = BodyData(3, fn["v"], fn["p"], fn["tv"], fn["tp"], None, fn["c"], fn["tc"])
    b = BodyData(3, verts, pose, tv, tp, None, clo, tc)
    assert verts.mean() > 0.3                                       # the inputs were not modified in place
    for k in ("vertices_train", "vertices_val", "vertices_test", "cond1_train", "cond1_val", "cond1_test", "cond2_train"):
        assert np.array_equal(getattr(a, k), getattr(b, k)) and getattr(a, k).dtype == np.float32
    mean, std = verts[:7].mean(0), verts[:7].std(0)                  # statistics of the TRAIN split only (:51-52)
    assert np.allclose(a.mean, mean) and np.allclose(a.std, std)
    assert np.allclose(a.vertices_val, (verts[7:] - mean) / std, atol=1e-5)
    assert np.allclose(a.vertices_test, (tv - mean) / std, atol=1e-5)
    # full 24-joint poses are reduced to the 14 clothing joints, the full ones kept (:93-98)
    assert a.cond1_train.shape == (7, 126) and a.cond1_test.shape == (3, 126) and a.cond1_train_full.shape == (7, 216)
    v, f = a.vec2mesh(a.vertices_train[0])
    assert np.allclose(v, verts[0], atol=1e-4) and f is None


def test_filter_cloth_pose_and_postprocess():
    from cape_b200 import demos
    p72 = np.arange(2 * 72, dtype=np.float64).reshape(2, 72)
    out = demos.filter_cloth_pose(p72)
    assert out.shape == (2, 42)
    assert np.array_equal(out[0, :3], p72[0, 3:6]) and np.array_equal(out[0, -3:], p72[0, 57:60])   # joints 1 and 19
    assert demos.filter_cloth_pose(np.zeros((3, 216))).shape == (3, 126)
    with pytest.raises(ValueError):
        demos.filter_cloth_pose(np.zeros((1, 10)))
    assert demos.useful_joints_idx == [1, 2, 3, 4, 5, 6, 9, 12, 13, 14, 16, 17, 18, 19]


def test_obj_roundtrip_and_demo_assets(tmp_path):
    from cape_b200 import demos, topology as T
    v, f = T.template_mesh()
    assert v.shape == (6890, 3) and f.shape == (13776, 3) and f.min() == 0 and f.max() == 6889
    fn = str(tmp_path / "m.obj")
    demos.write_obj(fn, v[:100], f[(f < 100).all(1)])
    v2, f2 = demos.read_obj(fn)
    assert np.allclose(v2, v[:100], atol=1e-7) and np.array_equal(f2, f[(f < 100).all(1)])
    rot, pose = T.demo_pose_params()
    assert rot.shape[1] == 216 and pose.shape[1] == 72
    mean, std = T.trainset_stats()
    keep = T.clothing_verts_idx()
    assert mean.shape == (6890, 3) and std.shape == (6890, 3) and keep.max() < 6890
    # the edge table is the upper triangle of the level-0 adjacency: every face edge is in it
    e = {tuple(x) for x in T.smpl_edges().tolist()}
    tri = f[:200]
    for a, b in ((0, 1), (1, 2), (0, 2)):
        assert all((min(x, y), max(x, y)) in e for x, y in zip(tri[:, a].tolist(), tri[:, b].tolist()))
