#!/usr/bin/env python
"""Per-row timings of SURVEY.md §8(a) on one GPU: every operator of the path at its
BASELINE.json configuration, CUDA events, inputs resident in HBM, against the
algorithmic bytes of SURVEY §8(d).  Secondary to bench.py (which measures the
headline C2 metric); output is a small table + JSON for profiles/.

    python tools/bench_rows.py [--iters 20] [--scene mixture|ground] [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import rd3_b200  # noqa: E402
from rd3_b200 import synthetic, voxel_layer  # noqa: E402

PEAK = 6551.4
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, iters, flush, graph=False):
    """Median / best CUDA-event time of fn().  graph=True: fn is captured once and the replay is timed -- what the
    kernels take without the Python / ctypes call in front of them (sub-0.2 ms calls are otherwise host-bound)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if graph:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
        fn = g.replay
        fn()
        torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.add_(1.0)                       # > L2: evict between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--scene", default="mixture")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    rows = collect(args.iters, args.scene)
    print("%-38s %-40s %10s %9s %9s %8s %10s" % ("row", "config", "alg MB", "ms(med)", "GB/s", "of HBM", "Munit/s"))
    for r in rows:
        print("%-38s %-40s %10.1f %9.3f %9.1f %7.1f%% %10.1f" % (r["row"], r["config"][:40], r["alg_MB"], r["ms_median"],
                                                                 r["GBps"], 100 * r["frac_of_hbm"], r["Munits_per_s"]))
    if args.json:
        json.dump(dict(peak_gbs=PEAK, scene=args.scene, rows=rows), open(args.json, "w"), indent=1)


def collect(iters=20, scene="mixture"):
    """-> list of row dicts (bench.py embeds them in its JSON line as `rows`)."""
    import types
    args = types.SimpleNamespace(iters=iters, scene=scene)
    dev = torch.device("cuda", torch.cuda.current_device())
    flush = torch.zeros(64 << 20, device=dev)            # 256 MB
    rows = []

    def add(name, cfg, units, unit_name, alg_bytes, ms):
        med, best = ms
        rows.append(dict(row=name, config=cfg, units=units, unit=unit_name, alg_MB=alg_bytes / 1e6,
                         ms_median=med, ms_best=best, GBps=alg_bytes / med / 1e6,
                         frac_of_hbm=alg_bytes / med / 1e6 / PEAK, Munits_per_s=units / med / 1e3))

    # ---- C2-shaped inputs ------------------------------------------------------------
    c2 = synthetic.CONFIGS["C2"]
    H, W = c2["hw"]
    B = 8
    b = synthetic.make_batch(list(range(B)), H, W, scene=args.scene)
    d = {k: v.to(dev) for k, v in b.items()}
    npix = 6 * H * W

    # a1 unprojection (module interface, padded output, no sync)
    fn = lambda: rd3_b200.unproject_padded(d["depth"], d["intrinsics"], d["cam2lidar"], max_depth=synthetic.MAX_DEPTH)
    pts, counts = fn()
    P = int(counts.sum())
    add("a1 unproject", "8 x 6x504x896, max_depth", B * npix, "pixel", B * npix * 4 + P * 12, timeit(fn, args.iters, flush))
    add("a1 unproject (graph replay)", "8 x 6x504x896, max_depth", B * npix, "pixel", B * npix * 4 + P * 12,
        timeit(fn, args.iters, flush, graph=True))
    thr = 1.4
    fn = lambda: rd3_b200.unproject_padded(d["depth"], d["intrinsics"], d["cam2lidar"], max_depth=synthetic.MAX_DEPTH,
                                           confs=d["conf"], conf_thresh=thr, sky_masks=d["sky"],
                                           range_filter=synthetic.FILTER_RANGE)
    pts2, counts2 = fn()
    P2 = int(counts2.sum())
    add("a1+a2 unproject+masks+range", "8 x 6x504x896", B * npix, "pixel", B * npix * 9 + P2 * 12,
        timeit(fn, args.iters, flush))
    add("a1+a2 (graph replay)", "8 x 6x504x896", B * npix, "pixel", B * npix * 9 + P2 * 12,
        timeit(fn, args.iters, flush, graph=True))

    # one frame's cloud for the point-array operators
    n0 = int(counts[0])
    cloud = pts[0, :n0].contiguous()
    N, C = cloud.shape

    # a3 dynamic voxelize
    vox_dyn = rd3_b200.Voxelization(c2["voxel_size"], c2["pcr"], -1)
    coors = vox_dyn(cloud)
    add("a3 dynamic_voxelize", "C3 grid 1440x1440x40, N=%d" % N, N, "point", N * 24, timeit(lambda: vox_dyn(cloud), args.iters, flush))
    add("a3 dynamic_voxelize (graph replay)", "C3 grid 1440x1440x40, N=%d" % N, N, "point", N * 24,
        timeit(lambda: vox_dyn(cloud), args.iters, flush, graph=True))

    # a4 hard voxelize (+a5 wrapper) C1/C2 grid and C4 pillars
    for name, cfg in (("C2", c2), ("C4", synthetic.CONFIGS["C4"])):
        K, mv = cfg["max_points"], cfg["max_voxels"][0]
        voxels = torch.empty((mv, K, C), device=dev)
        co = torch.empty((mv, 3), dtype=torch.int32, device=dev)
        nu = torch.empty((mv,), dtype=torch.int32, device=dev)
        mean = torch.empty((mv, C), device=dev)
        fnh = lambda: voxel_layer.hard_voxelize(cloud, voxels, co, nu, list(cfg["voxel_size"]), list(cfg["pcr"]), K, mv,
                                                voxel_mean=mean)
        M = fnh()
        add("a4+a9 hard_voxelize+mean (%s)" % name, "N=%d, M=%d, K=%d (incl. count D2H)" % (N, M, K), N, "point",
            N * 12 + M * (K * 12 + 16) + M * 12, timeit(fnh, args.iters, flush))
        vfe = rd3_b200.HardSimpleVFE(3)
        vv, nn = voxels[:M].contiguous(), nu[:M].contiguous()
        add("a9 HardSimpleVFE (%s)" % name, "M=%d, K=%d" % (M, K), M, "voxel", M * (K * 12 + 4 + 12),
            timeit(lambda: vfe(vv, nn, None), args.iters, flush))

    # a6/a7 DynamicScatter mean / max (C3)
    for avg in (True, False):
        ds = rd3_b200.DynamicScatter(c2["voxel_size"], c2["pcr"], avg)
        vf, vc = ds(cloud, coors)
        M = vf.shape[0]
        add("a6 DynamicScatter %s (C3)" % ("mean" if avg else "max"), "N=%d, M=%d (incl. M D2H)" % (N, M), N, "point",
            N * 28 + M * 28, timeit(lambda: ds(cloud, coors), args.iters, flush))
    # a7 batched mode (scatter_points.py:86-97): the B frames' clouds as one (sum N, 4) [b,z,y,x] call -- one launch
    # sequence for the batch instead of the reference's per-sample loop
    clouds = [pts[i, :int(counts[i])] for i in range(B)]
    bf = torch.cat(clouds).contiguous()
    bc = torch.cat([torch.nn.functional.pad(vox_dyn(c_), (1, 0), value=i) for i, c_ in enumerate(clouds)]).contiguous()
    for avg in (True, False):
        ds = rd3_b200.DynamicScatter(c2["voxel_size"], c2["pcr"], avg)
        vf, vc = ds(bf, bc)
        Nb, Mb = bf.shape[0], vf.shape[0]
        add("a7 DynamicScatter %s, batch of %d (C3)" % ("mean" if avg else "max", B), "sum N=%d, sum M=%d (incl. M D2H)" % (Nb, Mb),
            Nb, "point", Nb * 32 + Mb * 32, timeit(lambda: ds(bf, bc), args.iters, flush))
    # a8 backward
    f = cloud.clone().requires_grad_()
    ds = rd3_b200.DynamicScatter(c2["voxel_size"], c2["pcr"], True)
    vf, vc = ds(f, coors)
    g = torch.ones_like(vf)

    def bwd():
        f.grad = None
        vf.backward(g, retain_graph=True)
    add("a8 DynamicScatter backward mean", "N=%d, M=%d" % (N, vf.shape[0]), N, "point", N * 16 + vf.shape[0] * 16,
        timeit(bwd, args.iters, flush))

    # fused C2 / C4 (batched, 8 frames)
    for name, cfg in (("C2", c2), ("C4", synthetic.CONFIGS["C4"])):
        mod = rd3_b200.DepthToVoxels(cfg["voxel_size"], cfg["pcr"], cfg["max_points"], cfg["max_voxels"],
                                     max_depth=synthetic.MAX_DEPTH).to(dev).train()
        r = mod(d["depth"], d["intrinsics"], d["cam2lidar"])
        M = int(r["voxel_num"].sum())
        K = cfg["max_points"]
        add("fused depth->voxels (%s)" % name, "8 frames, M=%d" % M, B * npix, "pixel", B * npix * 4 + M * (K * 12 + 28),
            timeit(lambda: mod(d["depth"], d["intrinsics"], d["cam2lidar"]), args.iters, flush))

    # f2 packed sparse-encoder inputs (8 frames of the C2 result) and the voxel-free fused variant
    c2mod = rd3_b200.DepthToVoxels(c2["voxel_size"], c2["pcr"], c2["max_points"], c2["max_voxels"],
                                   max_depth=synthetic.MAX_DEPTH, with_voxels=False).to(dev).train()
    r = c2mod(d["depth"], d["intrinsics"], d["cam2lidar"])
    M = int(r["voxel_num"].sum())
    add("f2 pack_sparse_inputs", "8 frames, sum M=%d" % M, M, "voxel", M * (12 + 12 + 12 + 16),
        timeit(lambda: rd3_b200.pack_sparse_inputs(r, sync=False), args.iters, flush))
    add("f2 fused depth->encoder inputs (C2)", "8 frames, no voxel tensor", B * npix, "pixel",
        B * npix * 4 + M * 28 + M * 28,
        timeit(lambda: rd3_b200.pack_sparse_inputs(c2mod(d["depth"], d["intrinsics"], d["cam2lidar"]), sync=False),
               args.iters, flush))

    # f3 pillar decorations + scatter on the C4 pillars of one frame
    c4 = synthetic.CONFIGS["C4"]
    K4, mv4 = c4["max_points"], c4["max_voxels"][0]
    cloud5 = torch.cat([cloud, torch.rand(N, 2, device=dev)], dim=1).contiguous()
    v4 = torch.empty((mv4, K4, 5), device=dev)
    co4 = torch.empty((mv4, 3), dtype=torch.int32, device=dev)
    nu4 = torch.empty((mv4,), dtype=torch.int32, device=dev)
    M4 = voxel_layer.hard_voxelize(cloud5, v4, co4, nu4, list(c4["voxel_size"]), list(c4["pcr"]), K4, mv4)
    v4, nu4 = v4[:M4].contiguous(), nu4[:M4].contiguous()
    co4 = torch.nn.functional.pad(co4[:M4], (1, 0), value=0).contiguous()
    deco = rd3_b200.PillarDecorator(in_channels=5, voxel_size=c4["voxel_size"], point_cloud_range=c4["pcr"],
                                    legacy=False)
    add("f3 PillarFeatureNet decorations", "M=%d, K=%d, C 5->10" % (M4, K4), M4, "pillar",
        M4 * (K4 * 20 + 4 + 16 + K4 * 40), timeit(lambda: deco(v4, nu4, co4), args.iters, flush))
    pf = torch.randn(M4, 64, device=dev)
    sc = rd3_b200.PointPillarsScatter(64, (512, 512))
    add("f3 PointPillarsScatter", "M=%d, 64 ch -> 1x64x512x512" % M4, M4, "pillar",
        M4 * (256 + 16) + 64 * 512 * 512 * 4, timeit(lambda: sc(pf, co4, batch_size=1), args.iters, flush))

    # f4 confidence-percentile threshold (exact 3-pass radix select), 8 samples of 6x504x896
    add("f4 conf percentile (p=30, non-sky)", "8 x 6x504x896 conf + sky", B * npix, "pixel", B * npix * 5,
        timeit(lambda: rd3_b200.conf_threshold(d["conf"], d["sky"], 30.0), args.iters, flush))

    return rows


if __name__ == "__main__":
    main()
