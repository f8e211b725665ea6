"""Generate tests/golden/*.npz from the REFERENCE itself, in the build container.

Run here (needs /root/reference for oracle/_ref and torch for the restatement of
the Python-level reference code); the resulting small fixtures are committed and
travel to the GPU box, where /root/reference does not exist.

    python tests/golden/make_golden.py

Sources of truth:
  hard/dynamic voxelization : oracle/_ref = the reference's voxelization_cpu.cpp compiled unmodified
  unprojection              : torch-CPU ops of reconstruction_backbone.py:305-386 (oracle/torch_restatement.py)
  DynamicScatter            : torch-CPU ops of scatter_points_cuda.cu:183-239 (ibid.)
  pillar decorations/scatter: the reference's hard_voxelize (oracle/_ref) + torch-CPU ops of
                              pillar_encoder.py:104-146 and pillar_scatter.py:39-102 (ibid.)

    python tests/golden/make_golden.py pillar      # only (re)generate pillar.npz
    python tests/golden/make_golden.py real        # only (re)generate real_cloud.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from oracle import torch_restatement as tr  # noqa: E402
from rd3_b200 import synthetic  # noqa: E402
from test_oracle import _adversarial_points, kat_points  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def pillar(ref):
    """5. pillar encoders, gather side: 0.2 m pillars of a small two-sample cloud (5 features)"""
    vs, pcr, K = [0.2, 0.2, 8.0], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], 20
    g = torch.Generator().manual_seed(7)
    vox, num, coors = [], [], []
    for bi in range(2):
        f = synthetic.make_frame(200 + bi, 20, 36, scene="ground")
        pts = tr.backproject_depth_to_points(f["depth"][None], f["intrinsics"][None], f["cam2lidar"][None],
                                             max_depth=synthetic.MAX_DEPTH)[0]
        pts = torch.cat([pts, torch.rand(len(pts), 2, generator=g)], dim=1).contiguous()
        v, c, n = tr.voxelization_forward(ref, pts, vs, pcr, K, 400)
        vox.append(v); num.append(n)
        coors.append(torch.nn.functional.pad(c, (1, 0), value=bi))
    vox, num, coors = torch.cat(vox), torch.cat(num), torch.cat(coors)
    d = dict(voxels=vox.numpy(), num=num.numpy(), coors=coors.numpy(), voxel_size=vs, pcr=pcr)
    for legacy in (False, True):
        for dist in (False, True):
            d["deco_legacy%d_dist%d" % (legacy, dist)] = tr.pillar_feature_decorations(
                vox, num, coors, voxel_size=vs, point_cloud_range=pcr, with_distance=dist, legacy=legacy).numpy()
    feats = torch.randn(vox.shape[0], 8, generator=g)
    d["scatter_feats"] = feats.numpy()
    d["canvas"] = tr.point_pillars_scatter(feats, coors.long(), 8, 512, 512, batch_size=2).to_sparse().indices().numpy()
    d["canvas_values"] = tr.point_pillars_scatter(feats, coors.long(), 8, 512, 512, batch_size=2).to_sparse().values().numpy()
    np.savez_compressed(os.path.join(OUT, "pillar.npz"), **d)


def read_pcd_xyz(path):
    """xyz of a binary .pcd with FIELDS x y z rgb (the pseudo point clouds the reference ships under output/)"""
    raw = open(path, "rb").read()
    head, _, body = raw.partition(b"DATA binary\n")
    fields = [l.split()[1:] for l in head.decode().splitlines() if l.startswith("FIELDS")][0]
    npts = [int(l.split()[1]) for l in head.decode().splitlines() if l.startswith("POINTS")][0]
    a = np.frombuffer(body, dtype=np.float32, count=npts * len(fields)).reshape(npts, len(fields))
    return np.ascontiguousarray(a[:, :3])


def real_cloud(ref):
    """6. a REAL pseudo point cloud of the reference (output/sample_0_points.pcd, 40 000 points produced by its own
    DA3 -> unprojection -> FPS pipeline) through the reference's own CPU op: the C2 grid, a coarse grid
    with both truncations active, the pillar grid, and dynamic voxelization.  The padded voxel tensor is kept as per-voxel sums."""
    path = "/root/reference/output/sample_0_points.pcd"
    pts = read_pcd_xyz(path)
    d = dict(points=pts, source=os.path.basename(path))
    for tag, vs, pcr, mp, mv in (("c2", [0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3], 10, 120000),
                                 ("coarse", [0.5, 0.5, 0.5], [-54, -54, -5, 54, 54, 3], 2, 3000),
                                 ("c4", [0.2, 0.2, 8.0], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], 20, 30000)):
        v, c, n = tr.voxelization_forward(ref, torch.from_numpy(pts), vs, pcr, mp, mv)
        d[tag + "_cfg"] = np.array(vs + pcr + [mp, mv], dtype=np.float64)
        d[tag + "_coors"], d[tag + "_num"] = c.numpy(), n.numpy()
        d[tag + "_voxels_sum"] = v.sum(dim=1).numpy()
        d[tag + "_first"] = v[:, 0].numpy()                      # slot 0 of every voxel: its first point, bit for bit
    d["dyn_coors"] = tr.voxelization_forward(ref, torch.from_numpy(pts), [0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3],
                                             -1, -1).numpy()
    np.savez_compressed(os.path.join(OUT, "real_cloud.npz"), **d)


def main():
    ref = oracle.ref_voxel_layer()
    assert ref is not None, "build oracle/_ref first (python oracle/build_ref.py)"
    if sys.argv[1:] == ["real"]:
        real_cloud(ref)
        print("real_cloud.npz", os.path.getsize(os.path.join(OUT, "real_cloud.npz")))
        return
    if sys.argv[1:] == ["pillar"]:
        pillar(ref)
        print("pillar.npz", os.path.getsize(os.path.join(OUT, "pillar.npz")))
        return

    # 1. the reference's known-answer input through the reference's own op
    pts = torch.from_numpy(kat_points())
    v, c, n = tr.voxelization_forward(ref, pts, [0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1], 1000, 20000)
    np.savez_compressed(os.path.join(OUT, "kat_hard.npz"), points=pts.numpy(), coors=c.numpy(),
                        num=n.numpy(), voxels_sum=v.sum(dim=1).numpy())

    # 2. adversarial cloud, nuScenes grid, both truncations active
    vs, pcr, mp, mv = [0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3], 3, 2000
    p = _adversarial_points(20000, pcr, vs, seed=42)
    v, c, n = tr.voxelization_forward(ref, torch.from_numpy(p), vs, pcr, mp, mv)
    dc = tr.voxelization_forward(ref, torch.from_numpy(p), vs, pcr, -1, -1)
    np.savez_compressed(os.path.join(OUT, "adversarial_hard.npz"), points=p, voxel_size=vs, pcr=pcr,
                        max_points=mp, max_voxels=mv, voxels=v.numpy(), coors=c.numpy(), num=n.numpy(),
                        dyn_coors=dc.numpy())

    # 3. unprojection (+ masks) of 2 small frames
    b = synthetic.make_batch([100, 101], 28, 48)
    thr = tr.conf_threshold(b["conf"][0], b["sky"][0], synthetic.CONF_PERCENTILE)
    plain = tr.backproject_depth_to_points(b["depth"], b["intrinsics"], b["cam2lidar"],
                                           max_depth=synthetic.MAX_DEPTH)
    masked = tr.backproject_depth_to_points(b["depth"], b["intrinsics"], b["cam2lidar"],
                                            max_depth=synthetic.MAX_DEPTH, confs=b["conf"],
                                            conf_thresh=np.float64(thr), sky_masks=b["sky"])
    masked = [tr.filter_point_by_range(m, synthetic.FILTER_RANGE)[0] for m in masked]
    np.savez_compressed(os.path.join(OUT, "unproject.npz"), frame_ids=[100, 101], hw=[28, 48],
                        conf_thresh=thr, plain0=plain[0].numpy(), plain1=plain[1].numpy(),
                        masked0=masked[0].numpy(), masked1=masked[1].numpy())

    # 4. DynamicScatter
    g = torch.Generator().manual_seed(99)
    feats = torch.rand(5000, 4, generator=g) * 100 - 50
    coors = torch.randint(-1, 12, (5000, 3), generator=g, dtype=torch.int32)
    d = dict(feats=feats.numpy(), coors=coors.numpy())
    for red in ("sum", "mean", "max"):
        rf, rc, rm, rn = tr.dynamic_point_to_voxel_forward(feats, coors, red)
        d[red + "_feats"] = rf.numpy()
        d["voxel_coors"], d["map"], d["count"] = rc.numpy(), rm.numpy(), rn.numpy()
    np.savez_compressed(os.path.join(OUT, "dynamic_scatter.npz"), **d)
    pillar(ref)
    real_cloud(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
