"""Host-side logic that needs no GPU: sharding, the gloo (world_size 2) output gather,
threshold conversion, module attributes, the mmdet3d patch, the product/oracle firewall."""
import os
import re
import sys
import types

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import rd3_b200
from rd3_b200 import parallel, synthetic
from rd3_b200.backproject import f32_ceil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_f32_ceil_is_the_exact_fp64_threshold():
    g = np.random.default_rng(0)
    for x in list(g.random(2000) * 10) + [0.0, 1.5, -0.1, 1e-50, -1e-50, 3.4e38, 1.30000001]:
        f = f32_ceil(x)
        assert np.float32(f) == np.float64(f)              # representable in fp32
        assert np.float64(f) >= np.float64(x)
        below = np.nextafter(np.float32(f), np.float32(-np.inf))
        assert np.float64(below) < np.float64(x)
        # predicate equivalence on neighbouring fp32 values
        for c in (below, np.float32(f), np.nextafter(np.float32(f), np.float32(np.inf))):
            assert (np.float64(c) >= np.float64(x)) == (np.float32(c) >= np.float32(f))


def test_shard_range_partitions_the_batch():
    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 3, 4, 8):
            spans = [parallel.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == parallel.shard_sizes(n, w)
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def _fake_result(frame_ids, mv=6, k=2):
    """deterministic stand-in for DepthToVoxels output of the given global frames"""
    b = len(frame_ids)
    ids = torch.tensor(frame_ids, dtype=torch.float32).view(b, 1, 1, 1)
    return dict(voxels=ids.expand(b, mv, k, 3).clone() + torch.arange(mv).view(1, mv, 1, 1),
                coors=(ids.view(b, 1, 1).expand(b, mv, 3) * 10).to(torch.int32),
                num_points=torch.full((b, mv), k, dtype=torch.int32),
                voxel_mean=ids.view(b, 1, 1).expand(b, mv, 3).clone(),
                voxel_num=torch.tensor([1 + (i % mv) for i in frame_ids], dtype=torch.int32))


def _gloo_worker(rank, world, port, total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_range(total, world, rank)
    local = _fake_result(list(range(lo, hi)))
    full = parallel.gather_voxel_outputs(local, num_frames_total=total)
    ref = _fake_result(list(range(total)))
    ok = all(torch.equal(full[k], ref[k]) for k in ref)
    full2 = parallel.gather_voxel_outputs(local)              # sizes discovered by all_gather
    ok = ok and all(torch.equal(full2[k], ref[k]) for k in ref)
    try:                                   # packing is a CUDA kernel: CPU tensors must be refused
        parallel.to_sparse_encoder_inputs(full)
        ok = False
    except RuntimeError:
        pass
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _fake_flat_result(frame_ids, mv=7):
    """what DepthToVoxels(flat_outputs=True) returns, on the CPU: views of one int32 buffer"""
    r = _fake_result(frame_ids, mv=mv)
    B, MV, F = r["voxel_mean"].shape
    lay = parallel.flat_layout(B, MV, F)
    flat = torch.empty(lay[None], dtype=torch.int32)
    out = {"voxels": None, "flat": flat}
    for name in ("voxel_mean", "coors", "num_points", "voxel_num"):
        a, b, shape, dtype = lay[name]
        v = flat[a:b].view(dtype).view(shape)
        v.copy_(r[name])
        out[name] = v
    return out, MV


def _gloo_flat_worker(rank, world, port, total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.shard_range(total, world, rank)
    local, MV = _fake_flat_result(list(range(lo, hi)))
    views, _ = parallel.gather_flat_outputs(local, MV)
    ref = _fake_result(list(range(total)), mv=MV)
    ok = all(torch.equal(views[k], ref[k]) for k in ("voxel_mean", "coors", "num_points", "voxel_num"))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gloo_world2_flat_gather_is_one_collective_with_the_same_result():
    total = 6                                      # equal shards (3 + 3): the flat buffers have one size
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_flat_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
    # single process: a copy, same views
    local, MV = _fake_flat_result([0, 1, 2])
    views, work = parallel.gather_flat_outputs(local, MV)
    assert work is None and torch.equal(views["coors"], local["coors"]) and torch.equal(views["voxel_mean"], local["voxel_mean"])
    with pytest.raises(ValueError):
        parallel.gather_flat_outputs(dict(local, flat=local["flat"][:-1]), MV)


@pytest.mark.parametrize("total", [4, 5])
def test_gloo_world2_gather_matches_single_process(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + total
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_module_attributes_mirror_reference():
    v = rd3_b200.Voxelization([0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3], 10, (120000, 160000))
    assert v.grid_size.tolist() == [1440, 1440, 40]
    assert [int(x) for x in v.pcd_shape] == [1, 1440, 1440]
    assert v.max_voxels == (120000, 160000) and v.max_num_points == 10
    assert rd3_b200.Voxelization([0.5] * 3, [0, -40, -3, 70.4, 40, 1], 35).max_voxels == (20000, 20000)
    assert "max_num_points=10" in repr(v)
    d = rd3_b200.DynamicScatter([0.32, 0.32, 6], [-74.88, -74.88, -2, 74.88, 74.88, 4], True)
    assert d.average_points and "average_points=True" in repr(d)
    assert rd3_b200.HardSimpleVFE().num_features == 4


def test_synthetic_inputs_are_deterministic_and_shaped():
    a = synthetic.make_frame(3, 28, 48)
    b = synthetic.make_frame(3, 28, 48)
    for k in a:
        assert torch.equal(a[k], b[k], ) or (torch.isnan(a[k]) == torch.isnan(b[k])).all()
    d = a["depth"]
    assert d.shape == (6, 28, 48) and torch.isnan(d).any() and torch.isinf(d).any() and (d == 0).any()
    assert (a["conf"] >= 1).all() and a["sky"].dtype == torch.bool
    assert torch.equal(a["cam2lidar"][:, :3, 3], torch.zeros(6, 3))       # translation lives in ROW 3
    c = synthetic.make_frame(4, 28, 48)
    assert not torch.equal(torch.nan_to_num(c["depth"]), torch.nan_to_num(d))


def test_patch_mmdet3d_rebinds_standin_modules(monkeypatch):
    names = ["mmdet3d", "mmdet3d.ops", "mmdet3d.ops.voxel", "mmdet3d.ops.voxel.voxelize",
             "mmdet3d.ops.voxel.scatter_points", "mmdet3d.models", "mmdet3d.models.voxel_encoders",
             "mmdet3d.models.voxel_encoders.voxel_encoder"]
    mods = {n: types.ModuleType(n) for n in names}
    for n, m in mods.items():
        monkeypatch.setitem(sys.modules, n, m)
    vz, sp = mods["mmdet3d.ops.voxel.voxelize"], mods["mmdet3d.ops.voxel.scatter_points"]
    for a in ("hard_voxelize", "dynamic_voxelize", "Voxelization", "voxelization"):
        setattr(vz, a, None)
    for a in ("dynamic_point_to_voxel_forward", "dynamic_point_to_voxel_backward", "DynamicScatter", "dynamic_scatter"):
        setattr(sp, a, None)
    for a in ("Voxelization", "voxelization", "DynamicScatter", "dynamic_scatter"):
        setattr(mods["mmdet3d.ops"], a, None)

    class OldVoxelization:
        pass

    class OldDynamicScatter:
        pass
    vz.Voxelization, sp.DynamicScatter = OldVoxelization, OldDynamicScatter
    # a call site that did ``from mmdet3d.ops import DynamicScatter, Voxelization`` before the patch
    enc = mods["mmdet3d.models.voxel_encoders.voxel_encoder"]
    enc.DynamicScatter, enc.Voxelization = OldDynamicScatter, OldVoxelization

    class HardSimpleVFE:
        num_features = 4
    mods["mmdet3d.models.voxel_encoders.voxel_encoder"].HardSimpleVFE = HardSimpleVFE
    done = rd3_b200.patch_mmdet3d()
    assert vz.Voxelization is rd3_b200.Voxelization and vz.hard_voxelize is rd3_b200.voxel_layer.hard_voxelize
    assert sp.DynamicScatter is rd3_b200.DynamicScatter
    assert mods["mmdet3d.ops"].Voxelization is rd3_b200.Voxelization
    assert "mmdet3d.models.voxel_encoders.voxel_encoder.HardSimpleVFE.forward" in done
    assert enc.DynamicScatter is rd3_b200.DynamicScatter and enc.Voxelization is rd3_b200.Voxelization
    assert "mmdet3d.models.voxel_encoders.voxel_encoder.DynamicScatter" in done


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import or load it."""
    pkg = os.path.join(ROOT, "3d-reconstruction-detection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "librd3_oracle" not in src and "ref_voxel_layer" not in src, f
                assert "/root/reference" not in src, f


def test_committed_bench_line_keeps_the_contract():
    """profiles/r1_bench_default.json is a line bench.py printed on a B200: every key the bench
    contract names is there with a sane type (guards bench.py edits that cannot be run here)."""
    import json
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                        "r1_bench_default.json")
    j = json.load(open(path))
    for k, t in (("metric", str), ("value", float), ("unit", str), ("n_gpus", int), ("steps", int),
                 ("warmup", int), ("ms_per_step", float), ("higher_is_better", bool), ("scaling", str),
                 ("dtype", str), ("data", str), ("config", dict), ("clocks", dict), ("gpu_launches", int),
                 ("roofline", dict), ("e2e", dict), ("cpu_baseline", dict)):
        assert isinstance(j[k], t), k
    assert j["vs_baseline"] is None and j["scaling"] == "weak" and j["warmup"] >= 3 and j["gpu_launches"] > 0
    assert "workload" in j["config"] and "model" not in j["config"]
    r = j["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] is None or r["traffic"] > 0
    e = j["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < j["value"]
    c = j["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert set(j["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_raw_head_output_arguments():
    """DA3's raw outputs (output_processor.py:79-168): trailing singleton squeezed as a view, a floating-point
    sky tensor is a probability (thresholded in the kernels), bool / integer tensors are masks."""
    from rd3_b200.backproject import SKY_PROB_THRESH, make_params, sky_arg, squeeze_head
    d = torch.rand(2, 6, 4, 8, 1)
    v = squeeze_head(d)
    assert v.shape == (2, 6, 4, 8) and v.data_ptr() == d.data_ptr() and v.is_contiguous()
    assert squeeze_head(v) is v and squeeze_head(None) is None
    assert squeeze_head(torch.rand(2, 6, 4, 8, 3)).dim() == 5            # only a trailing 1 is a head dim
    prob = torch.rand(2, 6, 4, 8, 1)
    m, p = sky_arg(prob, "cpu")
    assert m is None and p.dtype == torch.float32 and p.shape == (2, 6, 4, 8) and p.data_ptr() == prob.data_ptr()
    m, p = sky_arg(prob.double(), "cpu")
    assert m is None and p.dtype == torch.float32
    for mask in (prob.squeeze(-1) >= 0.5, (prob.squeeze(-1) >= 0.5).to(torch.int32)):
        m, p = sky_arg(mask, "cpu")
        assert p is None and m.dtype == torch.uint8 and torch.equal(m.bool(), prob.squeeze(-1) >= 0.5)
    assert sky_arg(None, "cpu") == (None, None)
    assert SKY_PROB_THRESH == 0.5
    prm = make_params(2, 6, 4, 8, max_depth=100.0, sky_prob=p if p is not None else prob.squeeze(-1))
    assert prm.sky_prob == prob.data_ptr() and prm.sky_prob_thresh == 0.5
    assert make_params(2, 6, 4, 8).sky_prob is None
