"""Parity of the sm_100a path against the oracle.  Needs a CUDA device (-m gpu).

Bar: integer outputs (voxel coords, counts, point->voxel maps, voxel payload
copies) bit-exact; floating-point reductions within 1e-6 relative.
Nothing here reads /root/reference (it does not exist on the GPU box).
"""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import torch_restatement as tr
import rd3_b200
from rd3_b200 import synthetic, voxel_layer
from test_oracle import CASES, KAT_COORS, KAT_NUM, _adversarial_points, kat_points

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda:0"


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def gpu_hard(points, vs, pcr, mp, mv, mean_F=None, want_p2v=False):
    p = torch.from_numpy(np.ascontiguousarray(points, np.float32)).to(DEV)
    N, C = p.shape
    # poisoned buffers: every row < voxel_num must be fully overwritten
    voxels = torch.full((mv, mp, C), 777.0, device=DEV)
    coors = torch.full((mv, 3), -7, dtype=torch.int32, device=DEV)
    num = torch.full((mv,), -7, dtype=torch.int32, device=DEV)
    mean = torch.full((mv, mean_F), 777.0, device=DEV) if mean_F else None
    p2v = torch.full((N,), -7, dtype=torch.int32, device=DEV) if want_p2v else None
    n = voxel_layer.hard_voxelize(p, voxels, coors, num, vs, pcr, mp, mv, 3, True,
                                  voxel_mean=mean, point2voxel=p2v)
    out = [voxels[:n].cpu().numpy(), coors[:n].cpu().numpy(), num[:n].cpu().numpy()]
    # rows beyond voxel_num untouched
    assert (coors[n:] == -7).all() and (num[n:] == -7).all()
    if mean_F:
        out.append(mean[:n].cpu().numpy())
    if want_p2v:
        out.append(p2v.cpu().numpy())
    return out


def check_hard(points, vs, pcr, mp, mv):
    C = points.shape[1]
    v, c, n, mean, p2v = gpu_hard(points, vs, pcr, mp, mv, mean_F=min(C, 4), want_p2v=True)
    ov, oc, on, op2v = oracle.hard_voxelize(points, vs, pcr, mp, mv, return_point2voxel=True)
    assert len(c) == len(oc)
    assert np.array_equal(c, oc)
    assert np.array_equal(n, on)
    assert np.array_equal(bits(v), bits(ov))
    assert np.array_equal(p2v, op2v)
    om = oracle.hard_simple_vfe(ov, on, min(C, 4))
    assert np.array_equal(bits(mean), bits(om))          # same sequential order -> same bits
    o64 = oracle.hard_simple_vfe(ov, on, min(C, 4), f64=True)
    scale = np.abs(ov).max() if ov.size else 1.0
    assert np.allclose(mean, o64, rtol=1e-6, atol=1e-6 * scale)
    return len(c)


def frame_points(cfg, frame=0, scene="mixture"):
    H, W = synthetic.CONFIGS[cfg]["hw"]
    f = synthetic.make_frame(frame, H, W, scene=scene)
    pts = oracle.unproject(f["depth"].numpy(), f["intrinsics"].numpy(), f["cam2lidar"].numpy(),
                           max_depth=synthetic.MAX_DEPTH)
    return f, pts


# ----------------------------------------------------------------------------------------------
# library / boundary
# ----------------------------------------------------------------------------------------------
def test_no_cpu_path():
    with pytest.raises(RuntimeError):
        rd3_b200.Voxelization([0.5] * 3, [0, -40, -3, 70.4, 40, 1], 35)(torch.rand(10, 4))
    with pytest.raises(RuntimeError):
        voxel_layer.dynamic_point_to_voxel_forward(torch.rand(4, 3, device=DEV),
                                                   torch.zeros(4, 3, dtype=torch.int32, device=DEV), "min")


# ----------------------------------------------------------------------------------------------
# a3 dynamic_voxelize
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("vs,pcr,mp,mv", CASES)
def test_dynamic_voxelize_adversarial(vs, pcr, mp, mv):
    pts = _adversarial_points(40000, pcr, vs, seed=mp)
    for C in (3, 4, 5):
        p = pts[:, :C] if C <= 4 else np.concatenate([pts, pts[:, :1]], axis=1)
        p = np.ascontiguousarray(p)
        vox = rd3_b200.Voxelization(vs, pcr, -1)
        got = vox(torch.from_numpy(p).to(DEV)).cpu().numpy()
        assert np.array_equal(got, oracle.dynamic_voxelize(p, vs, pcr))


def test_dynamic_voxelize_c3_full_size():
    cfg = synthetic.CONFIGS["C3"]
    _, pts = frame_points("C3")
    got = rd3_b200.Voxelization(cfg["voxel_size"], cfg["pcr"], -1)(torch.from_numpy(pts).to(DEV))
    exp = oracle.dynamic_voxelize(pts, cfg["voxel_size"], cfg["pcr"])
    assert np.array_equal(got.cpu().numpy(), exp)
    assert (exp[:, 0] >= 0).sum() > 100000


# ----------------------------------------------------------------------------------------------
# a4 hard_voxelize (+ fused a9)
# ----------------------------------------------------------------------------------------------
def test_hard_known_answer():
    """mmdetection3d/tests/test_models/test_voxel_encoder/test_voxel_generator.py:7-22"""
    v, c, n = gpu_hard(kat_points(), [0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1], 1000, 20000)
    assert v.shape == (8, 1000, 4)
    assert np.array_equal(c, KAT_COORS)
    assert np.array_equal(n, KAT_NUM)
    g = np.load(os.path.join(GOLD, "kat_hard.npz"))
    assert np.allclose(v.sum(axis=1), g["voxels_sum"], rtol=1e-5)


def test_hard_golden_adversarial():
    g = np.load(os.path.join(GOLD, "adversarial_hard.npz"))
    vs, pcr = g["voxel_size"].tolist(), g["pcr"].tolist()
    v, c, n = gpu_hard(g["points"], vs, pcr, int(g["max_points"]), int(g["max_voxels"]))
    assert np.array_equal(c, g["coors"]) and np.array_equal(n, g["num"])
    assert np.array_equal(bits(v), bits(g["voxels"]))
    dyn = rd3_b200.Voxelization(vs, pcr, -1)(torch.from_numpy(g["points"]).to(DEV))
    assert np.array_equal(dyn.cpu().numpy(), g["dyn_coors"])


def test_hard_real_reference_cloud():
    """golden/real_cloud.npz: one of the pseudo point clouds the reference ships (output/sample_0_points.pcd) through
    the reference's own CPU op -- C2 grid, a coarse grid with both truncations active, the pillar grid, dynamic."""
    g = np.load(os.path.join(GOLD, "real_cloud.npz"))
    pts = g["points"]
    for tag in ("c2", "coarse", "c4"):
        cfg = g[tag + "_cfg"]
        vs, pcr, mp, mv = cfg[:3].tolist(), cfg[3:9].tolist(), int(cfg[9]), int(cfg[10])
        v, c, n = gpu_hard(pts, vs, pcr, mp, mv)
        assert np.array_equal(c, g[tag + "_coors"]) and np.array_equal(n, g[tag + "_num"]), tag
        assert np.array_equal(bits(v[:, 0]), bits(g[tag + "_first"])), tag
        assert np.allclose(v.sum(axis=1), g[tag + "_voxels_sum"], rtol=1e-5, atol=1e-4), tag
    dyn = rd3_b200.Voxelization([0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3], -1)(torch.from_numpy(pts).to(DEV))
    assert np.array_equal(dyn.cpu().numpy(), g["dyn_coors"])


@pytest.mark.parametrize("vs,pcr,mp,mv", CASES)
def test_hard_adversarial(vs, pcr, mp, mv):
    pts = _adversarial_points(60000, pcr, vs, seed=mp + 1)
    assert check_hard(pts, vs, pcr, mp, mv) > 0
    assert check_hard(np.ascontiguousarray(pts[:, :3]), vs, pcr, mp, mv) > 0


def test_hard_edge_cases():
    vs, pcr = [0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1]
    # empty input
    assert check_hard(np.zeros((0, 4), np.float32), vs, pcr, 5, 100) == 0
    # nothing in range
    assert check_hard(np.full((1000, 3), 1e6, np.float32), vs, pcr, 5, 100) == 0
    # one voxel, many points (max_points truncation under heavy contention), reversed chunks
    p = np.tile(np.array([[1.1, 1.1, 0.1, 0.0]], np.float32), (200000, 1))
    p[:, 3] = np.arange(200000)
    assert check_hard(p, vs, pcr, 7, 100) == 1
    # max_voxels = 1, max_points = 1
    pts = _adversarial_points(30000, pcr, vs, seed=5)
    assert check_hard(pts, vs, pcr, 1, 1) == 1
    # ragged: N not a multiple of anything
    assert check_hard(pts[:8193], vs, pcr, 3, 50) == 50
    assert check_hard(pts[:31], vs, pcr, 3, 50) > 0


@pytest.mark.parametrize("cfg,scene", [("C1", "mixture"), ("C2", "mixture"), ("C2", "ground"),
                                       ("C4", "mixture"), ("C4", "ground")])
def test_hard_full_size(cfg, scene):
    """BASELINE.json configs at full size: ~0.85 M / 2.7 M pixel frames."""
    c = synthetic.CONFIGS[cfg]
    _, pts = frame_points(cfg, scene=scene)
    for mv in c["max_voxels"]:
        m = check_hard(pts, list(c["voxel_size"]), list(c["pcr"]), c["max_points"], mv)
        assert m > 1000


def test_hard_deterministic_and_order_independent_of_scheduling():
    c = synthetic.CONFIGS["C4"]
    _, pts = frame_points("C1", scene="ground")
    a = gpu_hard(pts, list(c["voxel_size"]), list(c["pcr"]), c["max_points"], 30000)
    for _ in range(3):
        b = gpu_hard(pts, list(c["voxel_size"]), list(c["pcr"]), c["max_points"], 30000)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_voxelization_module_train_eval_max_voxels():
    c = synthetic.CONFIGS["C1"]
    _, pts = frame_points("C1")
    vox = rd3_b200.Voxelization(list(c["voxel_size"]), list(c["pcr"]), c["max_points"], (2000, 3000)).to(DEV)
    assert vox.grid_size.tolist() == [1440, 1440, 40]
    p = torch.from_numpy(pts).to(DEV)
    vox.train()
    v, co, n = vox(p)
    assert v.shape == (2000, 10, 3) and co.shape == (2000, 3) and n.shape == (2000,)
    vox.eval()
    v2, co2, n2 = vox(p)
    assert v2.shape[0] == 3000
    assert torch.equal(co2[:2000], co) and torch.equal(v2[:2000], v)
    assert co.dtype == torch.int32 and n.dtype == torch.int32


# ----------------------------------------------------------------------------------------------
# a9 HardSimpleVFE
# ----------------------------------------------------------------------------------------------
def test_hard_simple_vfe():
    g = np.random.default_rng(0)
    M, K, C = 24000, 10, 5
    n = g.integers(1, K + 1, size=M).astype(np.int32)
    v = (g.random((M, K, C)) * 10 + 20).astype(np.float32)
    for m in range(M):
        v[m, n[m]:] = 0
    vfe = rd3_b200.HardSimpleVFE(num_features=4)
    out = vfe(torch.from_numpy(v).to(DEV), torch.from_numpy(n).to(DEV), None)
    assert out.shape == (M, 4)                         # test_voxel_encoders.py:27-34 shape check
    assert np.array_equal(bits(out.cpu().numpy()), bits(oracle.hard_simple_vfe(v, n, 4)))
    # box-independent float check: fp64 anchor + the bound any fp32 summation order satisfies; the torch-CPU
    # restatement (whose reduction order depends on the host's vectorisation) must sit inside the same bound
    m64, tol = oracle.masked_mean_f64(v, n, 4)
    assert (np.abs(out.cpu().numpy() - m64) <= tol).all()
    ref = tr.hard_simple_vfe(torch.from_numpy(v), torch.from_numpy(n), 4).numpy()
    assert (np.abs(ref - m64) <= tol).all()
    assert tol.max() / 30.0 < 2e-6                      # the bound itself is inside the 1e-6-relative class


# ----------------------------------------------------------------------------------------------
# a1/a2 unprojection
# ----------------------------------------------------------------------------------------------
def test_unproject_golden():
    g = np.load(os.path.join(GOLD, "unproject.npz"))
    H, W = g["hw"].tolist()
    b = synthetic.make_batch(g["frame_ids"].tolist(), H, W)
    d = {k: v.to(DEV) for k, v in b.items()}
    pts, cols = rd3_b200.backproject_depth_to_points(d["depth"], d["intrinsics"], None, d["cam2lidar"],
                                                     max_depth=synthetic.MAX_DEPTH)
    assert cols == [None, None]
    for i in range(2):
        assert np.array_equal(bits(pts[i].cpu().numpy()), bits(g["plain%d" % i]))
    pts, _ = rd3_b200.backproject_depth_to_points(
        d["depth"], d["intrinsics"], None, d["cam2lidar"], max_depth=synthetic.MAX_DEPTH,
        multi_batch_confs=d["conf"], conf_thresh=float(g["conf_thresh"]),
        multi_batch_sky_masks=d["sky"], range_filter=synthetic.FILTER_RANGE)
    for i in range(2):
        assert np.array_equal(bits(pts[i].cpu().numpy()), bits(g["masked%d" % i]))


@pytest.mark.parametrize("hw,scene", [((280, 504), "mixture"), ((504, 896), "ground"), ((37, 53), "mixture")])
def test_unproject_vs_oracle(hw, scene):
    H, W = hw
    b = synthetic.make_batch([3, 4], H, W, scene=scene)
    d = {k: v.to(DEV) for k, v in b.items()}
    thr = tr.conf_threshold(b["conf"][0], b["sky"][0], synthetic.CONF_PERCENTILE)
    for masks in (False, True):
        kw = dict(max_depth=synthetic.MAX_DEPTH)
        okw = dict(max_depth=synthetic.MAX_DEPTH)
        if masks:
            kw.update(multi_batch_confs=d["conf"], conf_thresh=thr, multi_batch_sky_masks=d["sky"],
                      range_filter=synthetic.FILTER_RANGE)
        pts, _ = rd3_b200.backproject_depth_to_points(d["depth"], d["intrinsics"], None, d["cam2lidar"], **kw)
        for i in range(2):
            if masks:
                okw.update(conf=b["conf"][i].numpy(), conf_thresh=thr, sky=b["sky"][i].numpy(),
                           range_filter=synthetic.FILTER_RANGE)
            o = oracle.unproject(b["depth"][i].numpy(), b["intrinsics"][i].numpy(),
                                 b["cam2lidar"][i].numpy(), **okw)
            got = pts[i].cpu().numpy()
            assert got.shape == o.shape
            assert np.array_equal(bits(got), bits(o))


def test_unproject_mixin_and_colors():
    class Backbone(rd3_b200.DepthToPointsMixin):
        max_depth = synthetic.MAX_DEPTH
    H, W = 28, 48
    b = synthetic.make_batch([8], H, W)
    g = torch.Generator().manual_seed(0)
    imgs = torch.randint(0, 256, (1, 6, 3, H, W), generator=g).float()
    pts, cols = Backbone()._backproject_depth_to_points(b["depth"].to(DEV), b["intrinsics"].to(DEV),
                                                        imgs.to(DEV), b["cam2lidar"].to(DEV))
    o, pix = oracle.unproject(b["depth"][0].numpy(), b["intrinsics"][0].numpy(), b["cam2lidar"][0].numpy(),
                              max_depth=synthetic.MAX_DEPTH, return_pix=True)
    assert np.array_equal(bits(pts[0].cpu().numpy()), bits(o))
    flat = imgs[0].permute(0, 2, 3, 1).reshape(-1, 3).numpy()
    assert np.allclose(cols[0].cpu().numpy(), flat[pix] / 255.0, rtol=1e-6)
    # all-invalid depth -> empty (0,3) tensor
    z = torch.zeros(1, 6, H, W, device=DEV)
    pts, cols = Backbone()._backproject_depth_to_points(z, b["intrinsics"].to(DEV), None, b["cam2lidar"].to(DEV))
    assert tuple(pts[0].shape) == (0, 3)


# ----------------------------------------------------------------------------------------------
# fused depth -> voxels (C2, C4)
# ----------------------------------------------------------------------------------------------
def check_fused(cfg, frames, scene, masks, training):
    c = synthetic.CONFIGS[cfg]
    H, W = c["hw"]
    b = synthetic.make_batch(frames, H, W, scene=scene)
    d = {k: v.to(DEV) for k, v in b.items()}
    rf = synthetic.FILTER_RANGE if masks else None
    mod = rd3_b200.DepthToVoxels(c["voxel_size"], c["pcr"], c["max_points"], c["max_voxels"],
                                 max_depth=synthetic.MAX_DEPTH, range_filter=rf).to(DEV)
    mod.train(training)
    mv = c["max_voxels"][0 if training else 1]
    thr = tr.conf_threshold(b["conf"][0], b["sky"][0], synthetic.CONF_PERCENTILE) if masks else None
    r = mod(d["depth"], d["intrinsics"], d["cam2lidar"], confs=d["conf"] if masks else None,
            conf_thresh=thr, sky_masks=d["sky"] if masks else None)
    vn = r["voxel_num"].cpu().numpy()
    for i in range(len(frames)):
        okw = dict(max_depth=synthetic.MAX_DEPTH)
        if masks:
            okw.update(conf=b["conf"][i].numpy(), conf_thresh=thr, sky=b["sky"][i].numpy(), range_filter=rf)
        pts = oracle.unproject(b["depth"][i].numpy(), b["intrinsics"][i].numpy(), b["cam2lidar"][i].numpy(), **okw)
        ov, oc, on = oracle.hard_voxelize(pts, list(c["voxel_size"]), list(c["pcr"]), c["max_points"], mv)
        m = int(vn[i])
        assert m == len(oc)
        assert np.array_equal(r["coors"][i, :m].cpu().numpy(), oc)
        assert np.array_equal(r["num_points"][i, :m].cpu().numpy(), on)
        assert np.array_equal(bits(r["voxels"][i, :m].cpu().numpy()), bits(ov))
        om = oracle.hard_simple_vfe(ov, on, 3)
        assert np.array_equal(bits(r["voxel_mean"][i, :m].cpu().numpy()), bits(om))
    # batched sparse-encoder inputs == the slice + F.pad + cat tail of sparse_refinement.py:393-402
    feats, coors4, num, bs = mod.to_sparse_encoder_inputs(r, with_num_points=True)
    ef = torch.cat([r["voxel_mean"][i, :int(vn[i])] for i in range(len(frames))])
    ec = torch.cat([torch.nn.functional.pad(r["coors"][i, :int(vn[i])], (1, 0), value=i) for i in range(len(frames))])
    en = torch.cat([r["num_points"][i, :int(vn[i])] for i in range(len(frames))])
    assert bs == len(frames) and torch.equal(feats, ef) and torch.equal(coors4, ec) and torch.equal(num, en)
    f2, c2, offs = rd3_b200.pack_sparse_inputs(r, batch_offset=5, sync=False)
    assert offs.cpu().tolist() == [0] + np.cumsum(vn).tolist()
    assert torch.equal(c2[:len(ec), 0], ec[:, 0] + 5) and torch.equal(c2[:len(ec), 1:], ec[:, 1:])
    # the same call without the padded voxel tensor: coors / num / mean / counts identical
    keep = {k: v.clone() for k, v in r.items() if v is not None}
    lite = rd3_b200.DepthToVoxels(c["voxel_size"], c["pcr"], c["max_points"], c["max_voxels"],
                                  max_depth=synthetic.MAX_DEPTH, range_filter=rf, with_voxels=False).to(DEV)
    lite.train(training)
    r2 = lite(d["depth"], d["intrinsics"], d["cam2lidar"], confs=d["conf"] if masks else None,
              conf_thresh=thr, sky_masks=d["sky"] if masks else None)
    assert r2["voxels"] is None and torch.equal(r2["voxel_num"], keep["voxel_num"])
    for i in range(len(frames)):
        m = int(vn[i])
        assert torch.equal(r2["coors"][i, :m], keep["coors"][i, :m])
        assert torch.equal(r2["num_points"][i, :m], keep["num_points"][i, :m])
        assert np.array_equal(bits(r2["voxel_mean"][i, :m].cpu().numpy()), bits(keep["voxel_mean"][i, :m].cpu().numpy()))
    return vn


@pytest.mark.parametrize("cfg,scene,masks,training", [
    ("C2", "mixture", False, True), ("C2", "ground", True, False),
    ("C4", "mixture", True, True), ("C4", "ground", False, False),
    ("C1", "mixture", True, True)])
def test_fused_depth_to_voxels(cfg, scene, masks, training):
    vn = check_fused(cfg, [0, 1, 2], scene, masks, training)
    assert (vn > 1000).all()


# ----------------------------------------------------------------------------------------------
# a6/a7/a8 DynamicScatter
# ----------------------------------------------------------------------------------------------
def check_scatter(feats, coors, red, dims=None):
    f, c = torch.from_numpy(feats).to(DEV), torch.from_numpy(coors).to(DEV)
    vf, vc, p2v, cnt = voxel_layer.dynamic_point_to_voxel_forward(f, c, red, dims)
    of, oc, om, on = oracle.dynamic_scatter(feats, coors, red)
    assert np.array_equal(vc.cpu().numpy(), oc)
    assert np.array_equal(p2v.cpu().numpy(), om)
    assert np.array_equal(cnt.cpu().numpy(), on)
    got = vf.cpu().numpy()
    if red == "max":
        assert np.array_equal(bits(got), bits(of))
    else:
        # 1e-6 relative to the magnitude of the summands (the oracle is the fp64-accumulated value)
        scale = np.abs(feats).max()
        assert np.allclose(got, of, rtol=1e-6, atol=1e-6 * scale)
    return len(oc)


def test_dynamic_scatter_max_signed_zero():
    """-0.0 alone, with negatives, and with +0.0 in a voxel: the device fmaxf of the reference's reduceMax
    (scatter_points_cuda.cu:22-30) returns -0.0 / -0.0 / +0.0; a signed-int atomicMax on the bits would lose -0.0
    to the -inf initial value."""
    nz = np.float32(-0.0)
    feats = np.array([[nz, -1.0, nz], [nz, -2.0, 0.0], [-3.0, nz, -5.0], [nz, nz, nz], [-7.0, -0.5, nz]], np.float32)
    coors = np.array([[0, 0, 0], [0, 0, 0], [0, 0, 1], [2, 2, 2], [0, 0, 1]], np.int32)
    assert check_scatter(feats, coors, "max") == 3
    f, c = torch.from_numpy(feats).to(DEV), torch.from_numpy(coors).to(DEV)
    vf = voxel_layer.dynamic_point_to_voxel_forward(f, c, "max")[0].cpu().numpy()
    exp = np.array([[nz, -1.0, 0.0], [-3.0, nz, nz], [nz, nz, nz]], np.float32)
    assert np.array_equal(bits(vf), bits(exp))
    # max backward finds the arg-max of a -0.0 voxel (feats == voxel_feats)
    fg = f.clone().requires_grad_(True)
    out, _ = rd3_b200.DynamicScatter([1, 1, 1], [0, 0, 0, 4, 4, 4], False)(fg, c)
    out.sum().backward()
    assert float(fg.grad.sum()) == 9.0                 # every (voxel, feature) routed its gradient to one point


def test_dynamic_scatter_golden():
    g = np.load(os.path.join(GOLD, "dynamic_scatter.npz"))
    f, c = torch.from_numpy(g["feats"]).to(DEV), torch.from_numpy(g["coors"]).to(DEV)
    for red in ("sum", "mean", "max"):
        vf, vc, p2v, cnt = voxel_layer.dynamic_point_to_voxel_forward(f, c, red)
        assert np.array_equal(vc.cpu().numpy(), g["voxel_coors"])
        assert np.array_equal(p2v.cpu().numpy(), g["map"])
        assert np.array_equal(cnt.cpu().numpy(), g["count"])
        # the golden sums are torch-CPU fp32 index_add_ (order dependent): 1e-5 of the summand scale
        assert np.allclose(vf.cpu().numpy(), g[red + "_feats"], rtol=1e-5, atol=50 * 1e-5 if red == "sum" else 1e-4)


@pytest.mark.parametrize("red", ["sum", "mean", "max"])
def test_dynamic_scatter_reference_test_cases(red):
    """mmdetection3d/tests/test_models/test_voxel_encoder/test_dynamic_scatter.py:9-130 (forward part)"""
    g = torch.Generator().manual_seed(1)
    # empty input (:18-32)
    ef = torch.empty((0, 3), device=DEV)
    ec = torch.empty((0, 3), dtype=torch.int32, device=DEV)
    r = voxel_layer.dynamic_point_to_voxel_forward(ef, ec, red)
    assert r[0].shape == ef.shape and r[1].shape == ec.shape and r[2].numel() == 0 and r[3].numel() == 0
    # all points invalid (:35-50)
    feats = (torch.rand(200000, 3, generator=g) * 100 - 50).numpy()
    coors = torch.randint(-1, 0, (200000, 3), generator=g, dtype=torch.int32).numpy()
    assert check_scatter(feats, coors, red) == 0
    # random with / without negatives (:53-118)
    for low in (-1, 0):
        coors = torch.randint(low, 20, (200000, 3), generator=g, dtype=torch.int32).numpy()
        m = check_scatter(feats, coors, red)
        u = torch.from_numpy(coors).unique(dim=0, sorted=True)
        assert m == int((u.min(dim=-1).values >= 0).sum())
    # C = 1 and C = 7; hint too small -> transparent retry
    f7 = (torch.rand(5000, 7, generator=g) * 2 - 1).numpy()
    c7 = torch.randint(-2, 300, (5000, 3), generator=g, dtype=torch.int32).numpy()
    check_scatter(f7, c7, red)
    check_scatter(f7, c7, red, dims=[10, 10, 10])
    check_scatter(np.ascontiguousarray(f7[:, :1]), c7, red, dims=[300, 300, 300])


@pytest.mark.parametrize("red", ["mean", "max"])
def test_dynamic_scatter_c3_full_size(red):
    """C3: dynamic voxelization + DynamicScatter on the 1440x1440x40 grid, ~2.7 M points."""
    c = synthetic.CONFIGS["C3"]
    _, pts = frame_points("C3", scene="ground" if red == "max" else "mixture")
    p = torch.from_numpy(pts).to(DEV)
    coors = rd3_b200.Voxelization(c["voxel_size"], c["pcr"], -1)(p)
    ds = rd3_b200.DynamicScatter(c["voxel_size"], c["pcr"], red == "mean")
    vf, vc = ds(p, coors)
    of, oc, om, on = oracle.dynamic_scatter(pts, coors.cpu().numpy(), red)
    assert np.array_equal(vc.cpu().numpy(), oc)
    if red == "max":
        assert np.array_equal(bits(vf.cpu().numpy()), bits(of))
    else:
        assert np.allclose(vf.cpu().numpy(), of, rtol=1e-6, atol=1e-6 * 54)
    assert len(oc) > 100000


def test_dynamic_simple_vfe():
    """DynamicSimpleVFE (voxel_encoder.py:50-90) = no-grad DynamicScatter mean, single sample and batched (N, 4) coors."""
    g = torch.Generator().manual_seed(9)
    N = 20000
    feats = torch.rand(N, 4, generator=g) * 100 - 50
    coors = torch.randint(-1, 12, (N, 3), generator=g, dtype=torch.int32)
    vfe = rd3_b200.DynamicSimpleVFE([0.2, 0.2, 4], [0, -40, -3, 70.4, 40, 1]).to(DEV)
    f = feats.to(DEV).requires_grad_()
    vf, vc = vfe(f, coors.to(DEV))
    assert not vf.requires_grad                                   # @torch.no_grad() in the reference
    of, oc, _, _ = oracle.dynamic_scatter(feats.numpy(), coors.numpy(), "mean")
    assert np.array_equal(vc.cpu().numpy(), oc)
    assert np.allclose(vf.cpu().numpy(), of, rtol=1e-6, atol=5e-5)
    batch = torch.sort(torch.randint(0, 3, (N,), generator=g, dtype=torch.int32)).values
    coors4 = torch.cat([batch.view(-1, 1), coors], dim=1)
    vf4, vc4 = vfe(feats.to(DEV), coors4.to(DEV))
    rf, rc = tr.dynamic_scatter_batched(feats, coors4, "mean")
    assert torch.equal(vc4.cpu(), rc) and torch.allclose(vf4.cpu(), rf, rtol=1e-6, atol=5e-5)
    h, _ = vfe(feats.to(DEV).half(), coors.to(DEV))               # force_fp32(out_fp16=True)
    assert h.dtype == torch.half and torch.allclose(h.float().cpu(), torch.from_numpy(of).half().float(), atol=0.1)


def test_dynamic_scatter_batched_and_backward():
    g = torch.Generator().manual_seed(3)
    N = 30000
    feats = torch.rand(N, 4, generator=g) * 100 - 50
    coors = torch.randint(-1, 9, (N, 3), generator=g, dtype=torch.int32)
    batch = torch.sort(torch.randint(0, 3, (N,), generator=g, dtype=torch.int32)).values
    coors4 = torch.cat([batch.view(-1, 1), coors], dim=1)
    for avg in (True, False):
        red = "mean" if avg else "max"
        ds = rd3_b200.DynamicScatter([0.32, 0.32, 6], [-74.88, -74.88, -2, 74.88, 74.88, 4], avg)
        f = feats.clone().to(DEV).requires_grad_()
        vf, vc = ds(f, coors4.to(DEV))
        rf, rc = tr.dynamic_scatter_batched(feats, coors4, red)
        assert torch.equal(vc.cpu(), rc)
        assert torch.allclose(vf.detach().cpu(), rf, rtol=1e-6, atol=5e-5)
        # backward vs restated reference backward (scatter_points_cuda.cu:241-308)
        gout = torch.rand(vf.shape, generator=g)
        vf.backward(gout.to(DEV))
        exp = torch.zeros_like(feats)
        off = 0
        for i in range(3):
            m = coors4[:, 0] == i
            r = tr.dynamic_point_to_voxel_forward(feats[m].contiguous(), coors[m].contiguous(), red)
            M = r[0].shape[0]
            exp[m] = tr.dynamic_point_to_voxel_backward(gout[off:off + M], feats[m], r[0], r[2], r[3], red)
            off += M
        assert torch.allclose(f.grad.cpu(), exp, rtol=1e-6, atol=1e-7)
    # the batch is ONE launch sequence: the sample count is a hint that grows, never a host read per sample
    ds = rd3_b200.DynamicScatter([0.32, 0.32, 6], [-74.88, -74.88, -2, 74.88, 74.88, 4], True)
    assert ds._batch_hint == 1
    ds(feats.to(DEV), coors4.to(DEV))
    assert ds._batch_hint == 3
    # UNSORTED batch column: the reference takes batch_size from the LAST row (scatter_points.py:86) and never
    # looks at rows of a later sample; negative batch indices match no sample
    perm = torch.randperm(N, generator=g)
    c4 = coors4[perm].clone()
    c4[-1, 0] = 1                                        # batch_size = 2: the rows of sample 2 are ignored
    c4[:50, 0] = -1
    for red, avg in (("mean", True), ("max", False)):
        vf, vc = rd3_b200.DynamicScatter([0.32, 0.32, 6], [-74.88, -74.88, -2, 74.88, 74.88, 4], avg)(
            feats[perm].to(DEV), c4.to(DEV))
        rf, rc = tr.dynamic_scatter_batched(feats[perm], c4, red)
        assert int(vc[:, 0].max()) == 1 and torch.equal(vc.cpu(), rc)
        assert torch.allclose(vf.cpu(), rf, rtol=1e-6, atol=5e-5)
    # the 4-column C entry returns the point -> voxel map of the whole batch (ranks of the concatenated output)
    vf4, vc4, p2v4, cnt4 = voxel_layer.dynamic_point_to_voxel_forward(feats.to(DEV), coors4.to(DEV), "sum", [3, 9, 9, 9])
    assert int(cnt4.sum()) == int((p2v4 >= 0).sum()) and vc4.shape[1] == 4
    sel = (p2v4 >= 0).cpu()
    assert torch.equal(vc4.cpu()[p2v4.cpu()[sel].long()], coors4[sel])
    # all-invalid input -> zero grad (test_dynamic_scatter.py:35-50)
    f = feats.clone().to(DEV).requires_grad_()
    neg = torch.full((N, 3), -1, dtype=torch.int32, device=DEV)
    out, _ = rd3_b200.DynamicScatter([1, 1, 1], [0, 0, 0, 1, 1, 1], True)(f, neg)
    out.sum().backward()
    assert (f.grad == 0).all()


# ----------------------------------------------------------------------------------------------
# §8(f) next rows: pipeline steps of the plugin
# ----------------------------------------------------------------------------------------------
def test_voxel_downsample_and_range_filter(ref_layer):
    _, pts = frame_points("C1", scene="ground")
    pts = pts[:60000]
    g = torch.Generator().manual_seed(0)
    cols = torch.rand(len(pts), 3, generator=g)
    p = torch.from_numpy(pts)
    out = rd3_b200.FilterPointByRange(synthetic.FILTER_RANGE)({'points': p.to(DEV), 'colors': cols.to(DEV)})
    rp, ri = tr.filter_point_by_range(p, synthetic.FILTER_RANGE)
    assert torch.equal(out['points'].cpu(), rp) and torch.equal(out['indices'].cpu(), ri)
    fp, fc = out['points'], out['colors']
    ds = rd3_b200.VoxelDownsample(voxel_size=0.5, point_cloud_range=list(synthetic.FILTER_RANGE))
    r = ds({'points': fp, 'colors': fc})
    ec, ecol, eidx = tr.voxel_downsample(ref_layer, fp.cpu(), 0.5, list(synthetic.FILTER_RANGE), fc.cpu())
    assert r['points'].shape == ec.shape
    # centroid = mean of up to 100 points: fp64 anchor + summation-order bound for the kernel AND the torch restatement
    rv, _, rn = tr.voxelization_forward(ref_layer, fp.cpu().contiguous(), [0.5] * 3, list(synthetic.FILTER_RANGE), 100, 200000)
    c64, ctol = oracle.masked_mean_f64(rv.numpy(), rn.numpy())
    assert (np.abs(r['points'].cpu().numpy() - c64) <= ctol).all() and (np.abs(ec.numpy() - c64) <= ctol).all()
    # nearest-point colours: identical except where two points are equidistant to rounding
    same = (r['indices'].cpu() == eidx).float().mean().item()
    assert same > 0.999
    r2 = rd3_b200.VoxelDownsample(voxel_size=[0.5, 0.5, 0.5])({'points': fp})          # auto range, no colours
    ec2, _, _ = tr.voxel_downsample(ref_layer, fp.cpu(), 0.5, None)
    assert torch.allclose(r2['points'].cpu(), ec2, rtol=0, atol=float(ctol.max())) and r2['colors'] is None


# ----------------------------------------------------------------------------------------------
# randomized stress: odd shapes, tiny images, K = 1, small max_voxels, masks on/off, B > 1
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", list(range(8)))
def test_fused_randomized(seed):
    g = np.random.default_rng(1000 + seed)
    H, W = int(g.integers(5, 70)), int(g.integers(3, 140))
    ncam = int(g.integers(1, 7))
    B = int(g.integers(1, 4))
    K = int(g.choice([1, 2, 5, 10, 33]))
    mv = int(g.choice([1, 7, 300, 5000]))
    vs = [float(g.choice([0.075, 0.2, 0.5, 1.3])), float(g.choice([0.075, 0.2, 0.5])), float(g.choice([0.2, 1.0, 8.0]))]
    pcr = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
    masks = bool(g.integers(0, 2))
    scene = "ground" if g.integers(0, 2) else "mixture"
    frames = [synthetic.make_frame(int(g.integers(0, 10000)), H, W, num_cams=ncam, scene=scene) for _ in range(B)]
    b = {k: torch.stack([f[k] for f in frames]) for k in frames[0]}
    d = {k: v.to(DEV) for k, v in b.items()}
    rf = synthetic.FILTER_RANGE if masks else None
    mod = rd3_b200.DepthToVoxels(vs, pcr, K, mv, max_depth=synthetic.MAX_DEPTH, range_filter=rf).to(DEV)
    thr = 1.3 if masks else None
    r = mod(d["depth"], d["intrinsics"], d["cam2lidar"], confs=d["conf"] if masks else None, conf_thresh=thr,
            sky_masks=d["sky"] if masks else None)
    vn = r["voxel_num"].cpu().numpy()
    for i in range(B):
        okw = dict(max_depth=synthetic.MAX_DEPTH)
        if masks:
            okw.update(conf=b["conf"][i].numpy(), conf_thresh=thr, sky=b["sky"][i].numpy(), range_filter=rf)
        pts = oracle.unproject(b["depth"][i].numpy(), b["intrinsics"][i].numpy(), b["cam2lidar"][i].numpy(), **okw)
        ov, oc, on = oracle.hard_voxelize(pts, vs, pcr, K, mv)
        m = int(vn[i])
        assert m == len(oc), (seed, i, m, len(oc))
        assert np.array_equal(r["coors"][i, :m].cpu().numpy(), oc)
        assert np.array_equal(r["num_points"][i, :m].cpu().numpy(), on)
        assert np.array_equal(bits(r["voxels"][i, :m].cpu().numpy()), bits(ov))
        # the same cloud through the points API must give the same thing
        if len(pts):
            v2, c2, n2 = gpu_hard(pts, vs, pcr, K, mv)
            assert np.array_equal(c2, oc) and np.array_equal(n2, on) and np.array_equal(bits(v2), bits(ov))


@pytest.mark.parametrize("mv,masks", [(2500, False), (400, True)])
def test_fused_many_frames_per_lane(mv, masks):
    """The benchmark's batch shape in small: 24 frames = 12 per stream lane, frames of both scenes interleaved (so that
    the frames of one launch close in different insert rounds), every frame bit for bit against the oracle."""
    c = synthetic.CONFIGS["C1"]
    H, W, B = 60, 104, 24
    frames = [synthetic.make_frame(100 + i, H, W, scene="ground" if i % 3 == 1 else "mixture") for i in range(B)]
    b = {k: torch.stack([f[k] for f in frames]) for k in frames[0]}
    d = {k: v.to(DEV) for k, v in b.items()}
    rf = synthetic.FILTER_RANGE if masks else None
    thr = 1.4 if masks else None
    mod = rd3_b200.DepthToVoxels(c["voxel_size"], c["pcr"], c["max_points"], mv, max_depth=synthetic.MAX_DEPTH,
                                 range_filter=rf).to(DEV)
    r = mod(d["depth"], d["intrinsics"], d["cam2lidar"], confs=d["conf"] if masks else None, conf_thresh=thr,
            sky_masks=d["sky"] if masks else None)
    vn = r["voxel_num"].cpu().numpy()
    assert len(set(vn.tolist())) > 1 or vn[0] == mv
    for i in range(B):
        okw = dict(max_depth=synthetic.MAX_DEPTH)
        if masks:
            okw.update(conf=b["conf"][i].numpy(), conf_thresh=thr, sky=b["sky"][i].numpy(), range_filter=rf)
        pts = oracle.unproject(b["depth"][i].numpy(), b["intrinsics"][i].numpy(), b["cam2lidar"][i].numpy(), **okw)
        ov, oc, on = oracle.hard_voxelize(pts, list(c["voxel_size"]), list(c["pcr"]), c["max_points"], mv)
        m = int(vn[i])
        assert m == len(oc), (i, m, len(oc))
        assert np.array_equal(r["coors"][i, :m].cpu().numpy(), oc), i
        assert np.array_equal(r["num_points"][i, :m].cpu().numpy(), on), i
        assert np.array_equal(bits(r["voxels"][i, :m].cpu().numpy()), bits(ov)), i
        assert np.array_equal(bits(r["voxel_mean"][i, :m].cpu().numpy()), bits(oracle.hard_simple_vfe(ov, on, 3))), i


def test_cell_boundary_stress():
    """Depth planes chosen so that many unprojected points land (to rounding) ON voxel boundaries:
    the reciprocal fast path must hand every such pixel to the exact arithmetic."""
    H, W = 64, 128
    f = synthetic.make_frame(77, H, W, num_cams=2)
    K_, M_ = f["intrinsics"].clone(), f["cam2lidar"].clone()
    # axis-aligned camera looking along +x with zero translation: x_lidar = z_cam
    for n in range(2):
        M_[n] = torch.eye(4)
        M_[n, :3, :3] = torch.tensor([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
    g = torch.Generator().manual_seed(5)
    k = torch.randint(1, 700, (2, H, W), generator=g).float()
    depth = k * 0.075                      # exact multiples of the voxel size (and neighbours below)
    depth[:, ::2] = torch.nextafter(depth[:, ::2], torch.zeros(()))
    depth[:, 1::4] = torch.nextafter(depth[:, 1::4], torch.full((), 1e9))
    d = depth.unsqueeze(0).contiguous()
    vs, pcr = [0.075, 0.075, 0.2], [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
    mod = rd3_b200.DepthToVoxels(vs, pcr, 10, 20000, max_depth=100.0).to(DEV)
    r = mod(d.to(DEV), K_.unsqueeze(0).to(DEV), M_.unsqueeze(0).to(DEV))
    pts = oracle.unproject(depth.numpy(), K_.numpy(), M_.numpy(), max_depth=100.0)
    ov, oc, on = oracle.hard_voxelize(pts, vs, pcr, 10, 20000)
    m = int(r["voxel_num"][0])
    assert m == len(oc) and m > 500
    assert np.array_equal(r["coors"][0, :m].cpu().numpy(), oc)
    assert np.array_equal(r["num_points"][0, :m].cpu().numpy(), on)
    assert np.array_equal(bits(r["voxels"][0, :m].cpu().numpy()), bits(ov))


def test_hard_simple_vfe_backward():
    g = torch.Generator().manual_seed(2)
    M, K, C = 50, 6, 5
    n = torch.randint(1, K + 1, (M,), generator=g, dtype=torch.int32)
    x = torch.rand(M, K, C, generator=g)
    a = x.clone().to(DEV).requires_grad_()
    out = rd3_b200.HardSimpleVFE(4)(a, n.to(DEV), None)
    w = torch.rand(M, 4, generator=g)
    (out * w.to(DEV)).sum().backward()
    xr = x.clone().requires_grad_()
    (tr.hard_simple_vfe(xr, n, 4) * w).sum().backward()
    assert torch.allclose(a.grad.cpu(), xr.grad, rtol=1e-6, atol=1e-7)


def test_fused_outputs_are_fresh_unless_reuse_is_requested():
    """Like the reference ops, a second forward must not overwrite what the first returned (pred and GT
    voxelization in one training step); reuse_buffers=True is the explicit opt-in."""
    c = synthetic.CONFIGS["C1"]
    b0, b1 = synthetic.make_batch([0, 1], 56, 96), synthetic.make_batch([7, 8], 56, 96)
    d0, d1 = ({k: v.to(DEV) for k, v in b.items()} for b in (b0, b1))
    mod = rd3_b200.DepthToVoxels(c["voxel_size"], c["pcr"], c["max_points"], 3000, max_depth=synthetic.MAX_DEPTH).to(DEV)
    r0 = mod(d0["depth"], d0["intrinsics"], d0["cam2lidar"])
    keep = {k: v.clone() for k, v in r0.items()}
    r1 = mod(d1["depth"], d1["intrinsics"], d1["cam2lidar"])
    torch.cuda.synchronize()
    assert r0["coors"].data_ptr() != r1["coors"].data_ptr()
    for k in keep:
        assert torch.equal(r0[k], keep[k]), k
    assert not torch.equal(r0["voxel_num"], r1["voxel_num"]) or not torch.equal(r0["coors"], r1["coors"])
    p0 = rd3_b200.pack_sparse_inputs(r0)
    p0c = [t.clone() for t in p0[:2]]
    rd3_b200.pack_sparse_inputs(r1)
    assert all(torch.equal(a, b_) for a, b_ in zip(p0[:2], p0c))
    shared = rd3_b200.DepthToVoxels(c["voxel_size"], c["pcr"], c["max_points"], 3000, max_depth=synthetic.MAX_DEPTH,
                                    reuse_buffers=True).to(DEV)
    s0 = shared(d0["depth"], d0["intrinsics"], d0["cam2lidar"])
    s1 = shared(d1["depth"], d1["intrinsics"], d1["cam2lidar"])
    assert s0["coors"].data_ptr() == s1["coors"].data_ptr()


def test_nccl_gather_matches_single_process():
    """frames sharded by sample over 2 GPUs, outputs all-gathered with NCCL == the single-GPU result,
    bit for bit (unequal shards included).  Needs two GPUs; the gloo twin runs on the CPU (test_host_logic.py)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import subprocess
    import sys
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nccl_gather_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", worker],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


def test_cuda_graph_capture_of_fused_path():
    """The whole kernel sequence (incl. the internal stream lanes) is capturable: no host sync,
    no allocation inside DepthToVoxels.forward in steady state."""
    c = synthetic.CONFIGS["C1"]
    H, W = 56, 96
    b = synthetic.make_batch([0, 1, 2], H, W)
    d = {k: v.to(DEV) for k, v in b.items()}
    mod = rd3_b200.DepthToVoxels(c["voxel_size"], c["pcr"], c["max_points"], 3000, max_depth=synthetic.MAX_DEPTH,
                                 reuse_buffers=True).to(DEV)
    k_, m_ = d["intrinsics"].contiguous(), d["cam2lidar"].contiguous()
    ref = {k: v.clone() for k, v in mod(d["depth"], k_, m_).items()}        # warm-up allocates buffers
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        mod(d["depth"], k_, m_)                                               # workspace for stream s
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        out = mod(d["depth"], k_, m_)
    for _ in range(3):
        out["voxel_num"].zero_()
        graph.replay()
    torch.cuda.synchronize()
    vn = ref["voxel_num"].tolist()
    assert out["voxel_num"].tolist() == vn
    for i, m in enumerate(vn):
        assert torch.equal(out["coors"][i, :m], ref["coors"][i, :m])
        assert torch.equal(out["voxels"][i, :m], ref["voxels"][i, :m])


def test_flat_outputs_are_views_of_one_buffer():
    """DepthToVoxels(flat_outputs=True): mean / coors / num / voxel_num live in ONE int32 buffer (one collective moves a
    shard's encoder inputs); same values as the separate tensors."""
    c = synthetic.CONFIGS["C1"]
    b = synthetic.make_batch([0, 1, 2], 56, 96, scene="ground")
    d = {k: v.to(DEV) for k, v in b.items()}
    mk = lambda **kw: rd3_b200.DepthToVoxels(c["voxel_size"], c["pcr"], c["max_points"], 3000,
                                             max_depth=synthetic.MAX_DEPTH, **kw).to(DEV)
    a = mk(flat_outputs=True)(d["depth"], d["intrinsics"], d["cam2lidar"])
    r = mk()(d["depth"], d["intrinsics"], d["cam2lidar"])
    B, mv = 3, 3000
    flat = a["flat"]
    assert flat.dtype == torch.int32 and flat.numel() == B * mv * 7 + B
    assert a["voxel_mean"].data_ptr() == flat.data_ptr() and a["voxel_num"].data_ptr() == flat[B * mv * 7:].data_ptr()
    vn = r["voxel_num"].tolist()
    assert flat[B * mv * 7:].tolist() == vn
    for i, m in enumerate(vn):
        for k in ("coors", "num_points", "voxels", "voxel_mean"):
            assert torch.equal(a[k][i, :m], r[k][i, :m]), (i, k)
        assert torch.equal(flat[:B * mv * 3].view(torch.float32).view(B, mv, 3)[i, :m], r["voxel_mean"][i, :m])


def test_reused_buffers_alternating_scenes():
    """DepthToVoxels(reuse_buffers=True) over alternating dense / sparse scenes (and another max_voxels on the same
    module): nothing of an earlier call may survive in the scratch or the reused outputs."""
    c = synthetic.CONFIGS["C1"]
    H, W = 56, 96
    batches = [synthetic.make_batch([0, 1, 2], H, W, scene="ground"), synthetic.make_batch([3, 4, 5], H, W),
               synthetic.make_batch([6, 7, 8], H, W, scene="ground")]
    keep = rd3_b200.DepthToVoxels(c["voxel_size"], c["pcr"], c["max_points"], (3000, 700), max_depth=synthetic.MAX_DEPTH,
                                  reuse_buffers=True).to(DEV)
    fresh = rd3_b200.DepthToVoxels(c["voxel_size"], c["pcr"], c["max_points"], (3000, 700), max_depth=synthetic.MAX_DEPTH).to(DEV)
    for rep in range(2):
        for bi, b in enumerate(batches):
            for train in (True, False) if bi == 1 else (True,):
                keep.train(train); fresh.train(train)
                d = {k: v.to(DEV) for k, v in b.items()}
                a = keep(d["depth"], d["intrinsics"], d["cam2lidar"])
                r = fresh(d["depth"], d["intrinsics"], d["cam2lidar"])
                vn = r["voxel_num"].tolist()
                assert a["voxel_num"].tolist() == vn
                for i, m in enumerate(vn):
                    for k in ("coors", "num_points", "voxels", "voxel_mean"):
                        assert torch.equal(a[k][i, :m], r[k][i, :m]), (rep, bi, train, i, k)


# ----------------------------------------------------------------------------------------------
# pillar encoders, gather side
# ----------------------------------------------------------------------------------------------
def _pillar_inputs(seed, C=5, frames=(0, 1)):
    """C4 pillars of real voxelization output (+ random extra features), batched coors (b,z,y,x)."""
    c = synthetic.CONFIGS["C4"]
    g = torch.Generator().manual_seed(seed)
    vox, coors, num = [], [], []
    for bi, fid in enumerate(frames):
        f = synthetic.make_frame(fid, 60, 104, scene="ground")
        pts = oracle.unproject(f["depth"].numpy(), f["intrinsics"].numpy(), f["cam2lidar"].numpy(),
                               max_depth=synthetic.MAX_DEPTH)
        pts = np.concatenate([pts, torch.rand(len(pts), C - 3, generator=g).numpy()], axis=1).astype(np.float32)
        v, co, n = gpu_hard(pts, list(c["voxel_size"]), list(c["pcr"]), c["max_points"], 3000)
        vox.append(torch.from_numpy(v)); num.append(torch.from_numpy(n))
        coors.append(torch.nn.functional.pad(torch.from_numpy(co), (1, 0), value=bi))
    return torch.cat(vox), torch.cat(num), torch.cat(coors), c


@pytest.mark.parametrize("legacy,dist,cluster,center", [(False, False, True, True), (True, True, True, True),
                                                        (False, True, False, True), (True, False, True, False)])
def test_pillar_decorations(legacy, dist, cluster, center):
    vox, num, coors, c = _pillar_inputs(3)
    assert vox.shape[0] > 500 and int(num.max()) == c["max_points"] and int(num.min()) >= 1
    kw = dict(voxel_size=c["voxel_size"], point_cloud_range=c["pcr"], with_cluster_center=cluster,
              with_voxel_center=center, with_distance=dist, legacy=legacy)
    exp = tr.pillar_feature_decorations(vox, num, coors, **kw)
    mod = rd3_b200.PillarDecorator(in_channels=vox.shape[2], with_distance=dist, with_cluster_center=cluster,
                                   with_voxel_center=center, voxel_size=c["voxel_size"],
                                   point_cloud_range=c["pcr"], legacy=legacy)
    got = mod(vox.to(DEV), num.to(DEV), coors.to(DEV)).cpu()
    assert got.shape == exp.shape == (vox.shape[0], vox.shape[1], mod.out_channels)
    C = vox.shape[2]
    # raw / centre-offset columns: the same separately rounded fp32 operations -> same bits
    exact = list(range(C)) + ([C + 3 * cluster, C + 3 * cluster + 1] if center else [])
    assert np.array_equal(bits(got[:, :, exact].numpy()), bits(exp[:, :, exact].numpy()))
    # cluster offsets (sum order of the mean) and the norm: 1e-6 relative to the coordinate magnitude
    scale = float(vox[:, :, :3].abs().max())
    assert torch.allclose(got, exp, rtol=1e-6, atol=1e-6 * scale)
    # padded slots are exactly zero (sign included) wherever the reference's are
    pad = ~tr.get_paddings_indicator(num, vox.shape[1])
    assert np.array_equal(bits(got[pad].numpy()), bits(exp[pad].numpy()))


def test_point_pillars_scatter():
    vox, num, coors, c = _pillar_inputs(4)
    g = torch.Generator().manual_seed(9)
    feats = torch.randn(vox.shape[0], 64, generator=g)
    ny, nx = 512, 512
    mod = rd3_b200.PointPillarsScatter(64, (ny, nx))
    got = mod(feats.to(DEV), coors.to(DEV), batch_size=2).cpu()
    exp = tr.point_pillars_scatter(feats, coors.long(), 64, ny, nx, batch_size=2)
    assert got.shape == (2, 64, ny, nx) and np.array_equal(bits(got.numpy()), bits(exp.numpy()))
    one = coors[:, 0] == 1
    got1 = mod(feats[one].to(DEV), coors[one][:, 1:].contiguous().to(DEV))[0].cpu()      # forward_single, (z,y,x)
    exp1 = tr.point_pillars_scatter(feats[one], coors[one][:, 1:].long(), 64, ny, nx)[0]
    assert np.array_equal(bits(got1.numpy()), bits(exp1.numpy()))
    with pytest.raises(RuntimeError):
        mod(feats, coors, batch_size=2)                                                   # CPU tensors: no CPU path


def test_pillar_golden():
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pillar.npz"))
    vox, num, coors = (torch.from_numpy(d[k]).to(DEV) for k in ("voxels", "num", "coors"))
    vs, pcr = d["voxel_size"].tolist(), d["pcr"].tolist()
    scale = float(np.abs(d["voxels"][:, :, :3]).max())
    for legacy in (0, 1):
        for dist in (0, 1):
            got = rd3_b200.pillar_decorate(vox, num, coors, vs, pcr, True, True, bool(dist), bool(legacy)).cpu().numpy()
            exp = d["deco_legacy%d_dist%d" % (legacy, dist)]
            assert np.array_equal(bits(got[:, :, [0, 1, 2, 3, 4, 8, 9]]), bits(exp[:, :, [0, 1, 2, 3, 4, 8, 9]]))
            assert np.allclose(got, exp, rtol=1e-6, atol=1e-6 * scale)
    canvas = rd3_b200.PointPillarsScatter(8, (512, 512))(torch.from_numpy(d["scatter_feats"]).to(DEV), coors, 2)
    sp = canvas.cpu().to_sparse()
    assert np.array_equal(sp.indices().numpy(), d["canvas"]) and np.array_equal(sp.values().numpy(), d["canvas_values"])


def test_map_voxel_center_to_point():
    """DynamicVFE's cluster-centre step: dynamic voxelize -> batched DynamicScatter mean -> per-point
    gather of the voxel mean, against the reference's dense-canvas formulation."""
    vs, pcr = [0.5, 0.5, 0.5], [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
    pts, coors = [], []
    for bi in range(2):
        f = synthetic.make_frame(300 + bi, 40, 72, scene="ground")
        p = oracle.unproject(f["depth"].numpy(), f["intrinsics"].numpy(), f["cam2lidar"].numpy(),
                             max_depth=synthetic.MAX_DEPTH)
        c = oracle.dynamic_voxelize(p, vs, pcr)
        keep = c[:, 0] >= 0                                    # DynamicVFE sees in-range points only
        pts.append(torch.from_numpy(p[keep]))
        coors.append(torch.nn.functional.pad(torch.from_numpy(c[keep]), (1, 0), value=bi))
    pts, coors = torch.cat(pts).to(DEV), torch.cat(coors).to(DEV)
    vmean, vcoors = rd3_b200.DynamicScatter(vs, pcr, True)(pts, coors)
    assert vcoors.shape[1] == 4 and 1000 < vmean.shape[0] < pts.shape[0]
    got, idx = rd3_b200.map_voxel_center_to_point(coors, vmean, vcoors, return_index=True)
    exp = tr.map_voxel_center_to_point(coors.cpu().long(), vmean.cpu(), vcoors.cpu().long(), vs, pcr)
    assert np.array_equal(bits(got.cpu().numpy()), bits(exp.numpy()))
    assert torch.equal(vcoors[idx.long()], coors)              # every point found its own voxel
    # voxel rows in any order, duplicates and unknown voxels (-> row 0, the reference's zero canvas)
    perm = torch.randperm(vmean.shape[0], generator=torch.Generator().manual_seed(1)).to(DEV)
    got2 = rd3_b200.map_voxel_center_to_point(coors, vmean[perm].contiguous(), vcoors[perm].contiguous())
    assert torch.equal(got2, got)
    half = vmean.shape[0] // 2
    got3 = rd3_b200.map_voxel_center_to_point(coors, vmean[:half].contiguous(), vcoors[:half].contiguous())
    exp3 = tr.map_voxel_center_to_point(coors.cpu().long(), vmean[:half].cpu(), vcoors[:half].cpu().long(), vs, pcr)
    assert np.array_equal(bits(got3.cpu().numpy()), bits(exp3.numpy()))


# ----------------------------------------------------------------------------------------------
# the confidence-percentile threshold on the device
# ----------------------------------------------------------------------------------------------
def _conf_cases():
    g = torch.Generator().manual_seed(11)
    b = synthetic.make_batch([400, 401, 402], 70, 126)                    # 1 + Exp(1), 10 % sky
    yield "synthetic", b["conf"], b["sky"]
    q = (torch.rand(3, 6, 40, 64, generator=g) * 37).floor() / 8 + 1     # heavy ties
    yield "ties", q, torch.rand(3, 6, 40, 64, generator=g) < 0.5
    few = torch.ones(3, 6, 16, 24, dtype=torch.bool)
    few[0, 0, 0, :11] = False                                            # 11 non-sky pixels: still non-sky only
    few[1, 0, 0, :10] = False                                            # 10: falls back to every pixel
    yield "few", torch.randn(3, 6, 16, 24, generator=g) * 3, few         # negative values too
    yield "nosky", torch.rand(2, 1, 1, 7, generator=g), None
    withnan = torch.rand(2, 2, 8, 16, generator=g) + 1
    withnan[0, 1, 3, 5] = float("nan")                                   # np.percentile of data with a NaN: nan
    yield "nan", withnan, None


@pytest.mark.parametrize("pct", [30.0, 0.0, 100.0, 50.0, 99.9])
def test_conf_percentile_matches_numpy(pct):
    for name, conf, sky in _conf_cases():
        t32, t64 = rd3_b200.conf_threshold(conf.to(DEV), sky.to(DEV) if sky is not None else None, pct,
                                           numpy2=True, return_float64=True)
        o32, o64 = rd3_b200.conf_threshold(conf.to(DEV), sky.to(DEV) if sky is not None else None, pct,
                                           numpy2=False, return_float64=True)
        for i in range(conf.shape[0]):
            c, s = conf[i].numpy(), (sky[i].numpy() if sky is not None else None)
            px = c[~s] if (s is not None and (~s).sum() > 10) else c.flatten()
            exp = np.percentile(px, pct)                                  # the installed numpy (2.x): fp32 result
            assert exp.dtype == np.float32
            if np.isnan(exp):
                assert np.isnan(t32[i].item()) and np.isnan(t64[i].item()) and np.isnan(o64[i].item())
                continue
            assert bits(np.float32(t32[i].item())) == bits(exp), (name, i, pct, t32[i].item(), exp)
            assert t64[i].item() == float(exp)
            e1 = tr.conf_threshold_numpy1(conf[i], sky[i] if sky is not None else None, pct)
            assert o64[i].item() == e1, (name, i, pct, o64[i].item(), e1)
            assert bits(np.float32(o32[i].item())) == bits(np.float32(e1))


def test_conf_threshold_tensor_feeds_the_masks():
    """thresholds computed on the device and passed as a tensor == the host-float path per sample"""
    b = synthetic.make_batch([410, 411], 56, 96)
    d = {k: v.to(DEV) for k, v in b.items()}
    thr = rd3_b200.conf_threshold(d["conf"], d["sky"], synthetic.CONF_PERCENTILE)
    pts, cnt = rd3_b200.unproject_padded(d["depth"], d["intrinsics"], d["cam2lidar"], max_depth=synthetic.MAX_DEPTH,
                                         confs=d["conf"], conf_thresh=thr, sky_masks=d["sky"])
    mod = rd3_b200.DepthToVoxels([0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3], 10, 5000,
                                 max_depth=synthetic.MAX_DEPTH).to(DEV)
    r = mod(d["depth"], d["intrinsics"], d["cam2lidar"], confs=d["conf"], conf_thresh=thr, sky_masks=d["sky"])
    vnum = r["voxel_num"].clone()
    coors = r["coors"].clone()
    for i in range(2):
        t = tr.conf_threshold(b["conf"][i], b["sky"][i], synthetic.CONF_PERCENTILE)
        assert np.float32(t) == np.float32(thr[i].item())
        p1, c1 = rd3_b200.unproject_padded(d["depth"][i:i + 1], d["intrinsics"][i:i + 1], d["cam2lidar"][i:i + 1],
                                           max_depth=synthetic.MAX_DEPTH, confs=d["conf"][i:i + 1], conf_thresh=t,
                                           sky_masks=d["sky"][i:i + 1])
        n = int(c1[0])
        assert n == int(cnt[i]) and n > 1000 and torch.equal(p1[0, :n], pts[i, :n])
        r1 = mod(d["depth"][i:i + 1], d["intrinsics"][i:i + 1], d["cam2lidar"][i:i + 1], confs=d["conf"][i:i + 1],
                 conf_thresh=t, sky_masks=d["sky"][i:i + 1])
        m = int(r1["voxel_num"][0])
        assert m == int(vnum[i]) and torch.equal(r1["coors"][0, :m], coors[i, :m])


def test_raw_head_outputs_without_the_boolean_sky_tensor():
    """DA3's raw (B, N, H, W, 1) depth / conf / sky-probability outputs go straight into the
    kernels (sky iff prob >= 0.5, output_processor.py:152-168) == the restated squeeze + boolean mask path."""
    b = synthetic.make_batch([420, 421], 56, 96)
    g = torch.Generator().manual_seed(77)
    prob = torch.rand(b["sky"].shape, generator=g) * 0.5             # non-sky: [0, 0.5)
    prob = torch.where(b["sky"], 0.5 + prob, prob)                    # sky: [0.5, 1)
    flat = prob.view(-1)
    flat[5] = 0.5                                                     # exactly the threshold: sky
    flat[6] = float("nan")                                            # NaN >= 0.5 is False: not sky
    flat[7] = float(np.nextafter(np.float32(0.5), np.float32(0)))     # just below: not sky
    raw = dict(depth=b["depth"].unsqueeze(-1), conf=b["conf"].unsqueeze(-1), sky=prob.unsqueeze(-1))
    dep, conf, mask = tr.extract_head_outputs(raw["depth"], raw["conf"], raw["sky"])
    assert mask.view(-1)[5] and not mask.view(-1)[6] and not mask.view(-1)[7]
    d = {k: v.to(DEV) for k, v in b.items()}
    rawd = {k: v.to(DEV) for k, v in raw.items()}
    maskd = mask.to(DEV)
    # confidence threshold: raw probability == boolean mask == numpy on the restated mask
    t_raw = rd3_b200.conf_threshold(rawd["conf"], rawd["sky"], synthetic.CONF_PERCENTILE)
    t_msk = rd3_b200.conf_threshold(d["conf"], maskd, synthetic.CONF_PERCENTILE)
    assert torch.equal(t_raw, t_msk)
    for i in range(2):
        assert np.float32(tr.conf_threshold(conf[i], mask[i], synthetic.CONF_PERCENTILE)) == np.float32(t_raw[i].item())
    # unprojection and the fused path
    p_raw, c_raw = rd3_b200.unproject_padded(rawd["depth"], d["intrinsics"], d["cam2lidar"], max_depth=synthetic.MAX_DEPTH,
                                             confs=rawd["conf"], conf_thresh=t_raw, sky_masks=rawd["sky"])
    p_msk, c_msk = rd3_b200.unproject_padded(d["depth"], d["intrinsics"], d["cam2lidar"], max_depth=synthetic.MAX_DEPTH,
                                             confs=d["conf"], conf_thresh=t_msk, sky_masks=maskd)
    assert torch.equal(c_raw, c_msk)
    for i in range(2):
        n = int(c_raw[i])
        assert n > 1000 and torch.equal(p_raw[i, :n], p_msk[i, :n])
        o = oracle.unproject(dep[i].numpy(), b["intrinsics"][i].numpy(), b["cam2lidar"][i].numpy(),
                             max_depth=synthetic.MAX_DEPTH, conf=conf[i].numpy(),
                             conf_thresh=float(t_raw[i].item()), sky=mask[i].numpy())
        assert o.shape[0] == n and np.array_equal(bits(o), bits(p_raw[i, :n].cpu().numpy()))
    mod = rd3_b200.DepthToVoxels([0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3], 10, 5000,
                                 max_depth=synthetic.MAX_DEPTH).to(DEV)
    r = mod(rawd["depth"], d["intrinsics"], d["cam2lidar"], confs=rawd["conf"], conf_thresh=t_raw, sky_masks=rawd["sky"])
    r = {k: (v.clone() if v is not None else None) for k, v in r.items()}
    q = mod(d["depth"], d["intrinsics"], d["cam2lidar"], confs=d["conf"], conf_thresh=t_msk, sky_masks=maskd)
    assert torch.equal(r["voxel_num"], q["voxel_num"])
    for i in range(2):
        m = int(r["voxel_num"][i])
        assert m > 500
        for k in ("voxels", "coors", "num_points", "voxel_mean"):
            assert torch.equal(r[k][i, :m], q[k][i, :m]), k


def test_voxel_occupancy_and_dense_map():
    """GT-side occupancy: hard voxelization -> SoftVoxelOccupancyVFE -> dense (B, Z, Y, X) map."""
    vs, pcr, K = [0.6, 0.6, 0.8], [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0], 10      # occ grid 180 x 180 x 10
    Z, Y, X = 10, 180, 180
    vox, num, coors_list = [], [], []
    for bi in range(2):
        f = synthetic.make_frame(500 + bi, 48, 84, scene="ground")
        p = oracle.unproject(f["depth"].numpy(), f["intrinsics"].numpy(), f["cam2lidar"].numpy(),
                             max_depth=synthetic.MAX_DEPTH)
        v, c, n = gpu_hard(p, vs, pcr, K, 20000)
        vox.append(torch.from_numpy(v)); num.append(torch.from_numpy(n)); coors_list.append(torch.from_numpy(c))
    voxels, nums = torch.cat(vox), torch.cat(num)
    coors = torch.cat([torch.nn.functional.pad(c, (1, 0), value=i) for i, c in enumerate(coors_list)])
    assert voxels.shape[0] > 2000 and int(nums.max()) == K
    # The oracle is the reference formula in fp64 with a derived per-voxel bound (oracle.soft_voxel_occupancy_f64):
    # the variance cancels |xyz| ~ 50 m against |diff| ~ 0.2 m, so a 1-ulp change of the mean (torch-CPU's
    # sum order differs between hosts) moves p_occ by ~1e-5 -- a torch-CPU reduction is not a 1e-6 anchor.
    p64, tol = oracle.soft_voxel_occupancy_f64(voxels.numpy(), nums.numpy())
    got = rd3_b200.SoftVoxelOccupancyVFE()(voxels.to(DEV), nums.to(DEV), coors.to(DEV)).cpu()
    assert got.shape == p64.shape and got.dtype == torch.float32
    assert (np.abs(got.numpy().astype(np.float64) - p64) <= tol).all()
    exp = tr.soft_voxel_occupancy(voxels, nums)                      # the torch restatement obeys the same bound
    assert (np.abs(exp.numpy().astype(np.float64) - p64) <= tol).all()
    assert tol.max() < 5e-4 and np.median(tol) < 2e-5
    hard = rd3_b200.HardVoxelOccupancyVFE()(voxels.to(DEV), nums.to(DEV), coors.to(DEV)).cpu()
    assert torch.equal(hard, (nums > 0).float().view(-1, 1))
    occ, dense = rd3_b200.voxel_occupancy(voxels.to(DEV), nums.to(DEV), coors.to(DEV), dense_shape=(Z, Y, X),
                                          batch_size=2)
    exp_map = tr.occupancy_feature_map(occ.cpu(), coors_list, 2, Z, Y, X)
    assert torch.equal(dense.cpu(), exp_map) and torch.equal(occ.cpu(), got)
    assert int((dense != 0).sum()) == voxels.shape[0]
