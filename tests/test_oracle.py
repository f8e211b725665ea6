"""Pin the oracle: known-answer vector, the compiled reference op, torch restatement.

CPU only.  These tests are what allows tests/test_*_gpu.py to trust ``oracle``.
"""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import torch_restatement as tr
from rd3_b200 import synthetic


# --- known-answer test of the reference ------------------------------------------
# mmdetection3d/tests/test_models/test_voxel_encoder/test_voxel_generator.py:7-22
KAT_COORS = np.array([[7, 81, 1], [6, 81, 0], [7, 80, 1], [6, 81, 1],
                      [7, 81, 0], [6, 80, 1], [7, 80, 0], [6, 80, 0]])
KAT_NUM = np.array([120, 121, 127, 134, 115, 127, 125, 131])


def kat_points():
    np.random.seed(0)
    return np.random.rand(1000, 4).astype(np.float32)


def test_known_answer_c_oracle():
    v, c, n = oracle.hard_voxelize(kat_points(), [0.5, 0.5, 0.5],
                                   [0, -40, -3, 70.4, 40, 1], 1000, 20000)
    assert v.shape == (8, 1000, 4)
    assert np.all(c == KAT_COORS)
    assert np.all(n == KAT_NUM)


def test_known_answer_reference_op(ref_layer):
    pts = torch.from_numpy(kat_points())
    v, c, n = tr.voxelization_forward(ref_layer, pts, [0.5, 0.5, 0.5],
                                      [0, -40, -3, 70.4, 40, 1], 1000, 20000)
    assert tuple(v.shape) == (8, 1000, 4)
    assert np.all(c.numpy() == KAT_COORS)
    assert np.all(n.numpy() == KAT_NUM)


def _adversarial_points(n, pcr, vs, seed):
    """random cloud + exact boundaries, NaN/Inf/huge, duplicates."""
    g = np.random.default_rng(seed)
    lo, hi = np.array(pcr[:3], np.float32), np.array(pcr[3:], np.float32)
    span = hi - lo
    p = (lo - 0.1 * span + 1.2 * span * g.random((n, 3))).astype(np.float32)
    k = n // 8
    # points exactly on voxel boundaries
    idx = g.integers(0, 50, size=(k, 3)).astype(np.float32)
    p[:k] = lo + idx * np.array(vs, np.float32)
    p[k:k + 6] = [lo, hi, [lo[0], hi[1], lo[2]], np.nextafter(hi, lo), np.nextafter(lo, lo - 1), np.nextafter(hi, hi + 1)]
    p[k + 6] = [np.nan, 0, 0]
    p[k + 7] = [0, np.inf, 0]
    p[k + 8] = [0, 0, -np.inf]
    p[k + 9] = [3e38, 0, 0]
    p[k + 10] = [-3e38, 1e30, 0]
    p[k + 11] = [5e9, 0, 0]      # floor() beyond int32
    # clustered duplicates so max_points truncates
    p[-k:] = p[g.integers(0, 16, size=k)]
    feat = g.random((n, 1)).astype(np.float32)
    return np.concatenate([p, feat], axis=1)


CASES = [
    ([0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1], 35, 20000),
    ([0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3], 10, 120000),
    ([0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3], 3, 500),      # both truncations bite
    ([0.2, 0.2, 8.0], [-51.2, -51.2, -5, 51.2, 51.2, 3], 20, 30000),
    ([0.32, 0.32, 6.0], [-74.88, -74.88, -2, 74.88, 74.88, 4], 5, 2000),
]


@pytest.mark.parametrize("vs,pcr,mp,mv", CASES)
def test_c_oracle_matches_reference_op(ref_layer, vs, pcr, mp, mv):
    pts = _adversarial_points(40000, pcr, vs, seed=len(vs) + mp)
    v, c, n = oracle.hard_voxelize(pts, vs, pcr, mp, mv)
    rv, rc, rn = tr.voxelization_forward(ref_layer, torch.from_numpy(pts), vs, pcr, mp, mv)
    assert np.array_equal(c, rc.numpy())
    assert np.array_equal(n, rn.numpy())
    assert np.array_equal(v.view(np.uint32), rv.numpy().view(np.uint32))
    dc = oracle.dynamic_voxelize(pts, vs, pcr)
    rdc = tr.voxelization_forward(ref_layer, torch.from_numpy(pts), vs, pcr, -1, -1)
    assert np.array_equal(dc, rdc.numpy())
    # grid size mirrors voxelize.py:113-121
    g = torch.round((torch.tensor(pcr[3:], dtype=torch.float32) - torch.tensor(pcr[:3], dtype=torch.float32))
                    / torch.tensor(vs, dtype=torch.float32)).long().numpy()
    assert np.array_equal(oracle.grid_size(vs, pcr), g)


def test_c_oracle_on_a_real_reference_cloud():
    """golden/real_cloud.npz: one of the pseudo point clouds the reference ships (output/sample_0_points.pcd, 40 000
    points out of its own DA3 -> unprojection -> FPS pipeline) through the reference's own CPU op (oracle/_ref), made
    by tests/golden/make_golden.py.  The C restatement must reproduce it: coordinates, counts, first points bit for
    bit, slot sums."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "real_cloud.npz"))
    pts = g["points"]
    assert pts.shape == (40000, 3) and np.isfinite(pts).all()
    for tag in ("c2", "coarse", "c4"):
        cfg = g[tag + "_cfg"]
        vs, pcr, mp, mv = cfg[:3].tolist(), cfg[3:9].tolist(), int(cfg[9]), int(cfg[10])
        v, c, n = oracle.hard_voxelize(pts, vs, pcr, mp, mv)
        assert np.array_equal(c, g[tag + "_coors"]) and np.array_equal(n, g[tag + "_num"]), tag
        assert np.array_equal(v[:, 0].view(np.uint32), np.ascontiguousarray(g[tag + "_first"]).view(np.uint32)), tag
        assert np.allclose(v.sum(axis=1), g[tag + "_voxels_sum"], rtol=1e-5, atol=1e-4), tag
    assert g["coarse_num"].max() == 2 and len(g["coarse_num"]) == 3000        # both truncations were active
    dc = oracle.dynamic_voxelize(pts, [0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3])
    assert np.array_equal(dc, g["dyn_coors"])
    assert (g["dyn_coors"][:, 0] < 0).any()                                 # the cloud has points outside the grid (z up to 6 m)


def test_hard_consistent_with_dynamic():
    """test_voxelize.py:50-59 logic: every hard voxel's points are the points dynamic
    voxelization maps to that coordinate, in point order."""
    vs, pcr = [0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1]
    pts = _adversarial_points(5000, pcr, vs, seed=3)
    v, c, n = oracle.hard_voxelize(pts, vs, pcr, 1000, 20000)
    dc = oracle.dynamic_voxelize(pts, vs, pcr)
    for i in range(len(c)):
        idx = np.all(dc == c[i], axis=1)
        assert idx.sum() == n[i] > 0
        assert np.array_equal(pts[idx], v[i][:n[i]])


def test_numba_voxel_generator_cross_check():
    """Second independent reference: mmdet3d/core/voxel/voxel_generator.py:137-208
    loaded standalone (numba).  Skipped where /root/reference is absent."""
    import importlib.util
    import os
    p = "/root/reference/mmdetection3d/mmdet3d/core/voxel/voxel_generator.py"
    if not os.path.exists(p):
        pytest.skip("reference tree not present")
    pytest.importorskip("numba")
    spec = importlib.util.spec_from_file_location("ref_voxel_generator", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    vs, pcr = [0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1]
    pts = _adversarial_points(20000, pcr, vs, seed=11)
    pts = pts[np.isfinite(pts).all(axis=1) & (np.abs(pts) < 1e6).all(axis=1)]
    gen = mod.VoxelGenerator(vs, pcr, 35, 20000)
    gv, gc, gn = gen.generate(pts)
    v, c, n = oracle.hard_voxelize(pts, vs, pcr, 35, 20000)
    assert np.array_equal(gc, c)
    assert np.array_equal(gn, n)
    assert np.array_equal(gv, v)


def test_vfe_matches_torch():
    g = np.random.default_rng(0)
    v = (g.random((5000, 10, 4)) * 100 - 50).astype(np.float32)
    n = g.integers(1, 11, size=5000).astype(np.int32)
    for m in range(5000):
        v[m, n[m]:] = 0
    a = oracle.hard_simple_vfe(v, n, 3)
    b = tr.hard_simple_vfe(torch.from_numpy(v), torch.from_numpy(n), 3).numpy()
    # torch's reduction order over dim=1 is not the sequential one, so fp32 results agree to
    # rounding only: 1e-6 relative to the magnitude of the summands (|v| <= 50 here).
    c = oracle.hard_simple_vfe(v, n, 3, f64=True)
    assert np.allclose(a, b, rtol=1e-6, atol=50e-6)
    assert np.allclose(a, c, rtol=1e-6, atol=50e-6)
    assert np.allclose(b, c, rtol=1e-6, atol=50e-6)


@pytest.mark.parametrize("scene", ["mixture", "ground"])
def test_unproject_matches_torch_restatement(scene):
    """Oracle's explicit FMA chain == torch-CPU `pts @ R.T + t` on this machine
    (SURVEY Appendix B4).  >=100 points per camera (torch takes another path below)."""
    H, W = 56, 96
    b = synthetic.make_batch([0, 1], H, W, scene=scene)
    ref = tr.backproject_depth_to_points(b["depth"], b["intrinsics"], b["cam2lidar"],
                                         max_depth=synthetic.MAX_DEPTH)
    for i in range(2):
        o = oracle.unproject(b["depth"][i].numpy(), b["intrinsics"][i].numpy(),
                             b["cam2lidar"][i].numpy(), max_depth=synthetic.MAX_DEPTH)
        r = ref[i].numpy()
        assert o.shape == r.shape
        assert np.array_equal(o.view(np.uint32), r.view(np.uint32))


def test_unproject_masks_and_range_filter():
    H, W = 56, 96
    b = synthetic.make_batch([5], H, W)
    thr = tr.conf_threshold(b["conf"][0], b["sky"][0], synthetic.CONF_PERCENTILE)
    ref = tr.backproject_depth_to_points(b["depth"], b["intrinsics"], b["cam2lidar"],
                                         max_depth=synthetic.MAX_DEPTH, confs=b["conf"],
                                         conf_thresh=np.float32(thr), sky_masks=b["sky"])[0]
    ref_f, _ = tr.filter_point_by_range(ref, synthetic.FILTER_RANGE)
    o = oracle.unproject(b["depth"][0].numpy(), b["intrinsics"][0].numpy(), b["cam2lidar"][0].numpy(),
                         max_depth=synthetic.MAX_DEPTH, conf=b["conf"][0].numpy(), conf_thresh=thr,
                         sky=b["sky"][0].numpy(), range_filter=synthetic.FILTER_RANGE)
    assert np.array_equal(o.view(np.uint32), ref_f.numpy().view(np.uint32))
    assert 0 < len(o) < 6 * H * W


@pytest.mark.parametrize("reduce_type", ["sum", "mean", "max"])
def test_dynamic_scatter_c_vs_torch(reduce_type):
    g = torch.Generator().manual_seed(7)
    feats = torch.rand(20000, 4, generator=g) * 100 - 50
    coors = torch.randint(-1, 20, (20000, 3), generator=g, dtype=torch.int32)
    rf, rc, rm, rn = tr.dynamic_point_to_voxel_forward(feats, coors, reduce_type)
    of, oc, om, on = oracle.dynamic_scatter(feats.numpy(), coors.numpy(), reduce_type)
    assert np.array_equal(oc, rc.numpy())
    assert np.array_equal(om, rm.numpy())
    assert np.array_equal(on, rn.numpy())
    if reduce_type == "max":
        assert np.array_equal(of, rf.numpy())
    else:
        assert np.allclose(of, rf.numpy(), rtol=1e-5, atol=1e-4)
    # test_dynamic_scatter.py:57-58,80: coors == unique(dim=0) minus negatives
    u = coors.unique(dim=0, sorted=True)
    u = u[u.min(dim=-1).values >= 0]
    assert np.array_equal(oc, u.numpy())


def test_pillar_restatement_matches_golden():
    """oracle/torch_restatement.py pillar functions vs tests/golden/pillar.npz (generated from the
    reference's own hard_voxelize output, tests/golden/make_golden.py)."""
    from oracle import torch_restatement as tr
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pillar.npz"))
    vox, num, coors = torch.from_numpy(d["voxels"]), torch.from_numpy(d["num"]), torch.from_numpy(d["coors"])
    vs, pcr = d["voxel_size"].tolist(), d["pcr"].tolist()
    for legacy in (0, 1):
        for dist in (0, 1):
            got = tr.pillar_feature_decorations(vox, num, coors, voxel_size=vs, point_cloud_range=pcr,
                                                with_distance=bool(dist), legacy=bool(legacy))
            assert np.array_equal(got.numpy().view(np.uint32), d["deco_legacy%d_dist%d" % (legacy, dist)].view(np.uint32))
    canvas = tr.point_pillars_scatter(torch.from_numpy(d["scatter_feats"]), coors.long(), 8, 512, 512, batch_size=2)
    sp = canvas.to_sparse()
    assert np.array_equal(sp.indices().numpy(), d["canvas"]) and np.array_equal(sp.values().numpy(), d["canvas_values"])
    # legacy=True really aliases: the raw x column equals the centre-offset column
    leg = d["deco_legacy1_dist0"]
    assert np.array_equal(leg[:, :, 0], leg[:, :, 8]) and not np.array_equal(d["deco_legacy0_dist0"][:, :, 0], leg[:, :, 0])


def test_fp64_anchors_bound_the_torch_restatement():
    """The float parity tests compare against fp64 anchors with DERIVED bounds (oracle.soft_voxel_occupancy_f64,
    oracle.masked_mean_f64), not against a torch-CPU reduction whose order depends on the host: the
    restatement itself has to sit inside those bounds, with room to spare, on this machine too."""
    import torch
    from oracle import torch_restatement as tr
    from rd3_b200 import synthetic
    f = synthetic.make_frame(500, 48, 84, scene="ground")
    p = oracle.unproject(f["depth"].numpy(), f["intrinsics"].numpy(), f["cam2lidar"].numpy(), max_depth=synthetic.MAX_DEPTH)
    v, c, n = oracle.hard_voxelize(p, [0.6, 0.6, 0.8], [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0], 10, 20000)
    assert len(n) > 2000 and int(n.max()) == 10
    p64, tol = oracle.soft_voxel_occupancy_f64(v, n)
    exp = tr.soft_voxel_occupancy(torch.from_numpy(v), torch.from_numpy(n)).numpy()
    assert (np.abs(exp - p64) <= tol).all() and tol.max() < 5e-4
    m64, mt = oracle.masked_mean_f64(v, n)
    ref = tr.hard_simple_vfe(torch.from_numpy(v), torch.from_numpy(n), 3).numpy()
    assert (np.abs(ref - m64) <= mt).all()
    assert (np.abs(oracle.hard_simple_vfe(v, n, 3) - m64) <= mt).all()
