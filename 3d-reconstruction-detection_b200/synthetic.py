"""Deterministic synthetic nuScenes-shaped inputs.

Everything is generated on the CPU from ``torch.Generator().manual_seed(1234 +
frame_idx)`` so that the CPU oracle and the GPU path see identical bits.
"""
import math

import torch

# BASELINE.json configs (SURVEY.md §8 shorthand)
CONFIGS = {
    # ResDet3D_nuscenes_mini_config.py:9,14,237-258
    "C1": dict(hw=(280, 504), voxel_size=(0.075, 0.075, 0.2),
               pcr=(-54.0, -54.0, -5.0, 54.0, 54.0, 3.0), max_points=10,
               max_voxels=(120000, 160000)),
    "C2": dict(hw=(504, 896), voxel_size=(0.075, 0.075, 0.2),
               pcr=(-54.0, -54.0, -5.0, 54.0, 54.0, 3.0), max_points=10,
               max_voxels=(120000, 160000)),
    "C3": dict(hw=(504, 896), voxel_size=(0.075, 0.075, 0.2),
               pcr=(-54.0, -54.0, -5.0, 54.0, 54.0, 3.0), max_points=-1,
               max_voxels=(-1, -1)),
    # configs/_base_/models/centerpoint_02pillar_second_secfpn_nus.py:1-5
    "C4": dict(hw=(504, 896), voxel_size=(0.2, 0.2, 8.0),
               pcr=(-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), max_points=20,
               max_voxels=(30000, 40000)),
}
MAX_DEPTH = 100.0          # tools/inference_nuscenes.py max-depth mask
CONF_PERCENTILE = 30.0     # ResDet3D_nuscenes_mini_config.py:219
FILTER_RANGE = (-54.0, -54.0, -5.0, 54.0, 54.0, 6.0)  # mini_config :142-144
YAWS_DEG = (0.0, -55.0, 55.0, 180.0, 110.0, -110.0)


def _rot(axis, deg):
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    if axis == "z":
        return torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
    if axis == "y":
        return torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], dtype=torch.float64)
    return torch.tensor([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]], dtype=torch.float64)


def make_calibration(num_cams, H, W, gen):
    """intrinsics (num_cams,3,3) and cam2lidar (num_cams,4,4) in the reference's
    layout: rotation in M[:3,:3], translation in ROW 3 (reconstruction_backbone.py:370)."""
    K = torch.zeros(num_cams, 3, 3, dtype=torch.float32)
    M = torch.zeros(num_cams, 4, 4, dtype=torch.float32)
    f = 1266.0 * W / 1600.0
    # camera (x right, y down, z forward) -> lidar (x forward, y left, z up)
    swap = torch.tensor([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]], dtype=torch.float64)
    jit = (torch.rand(num_cams, 4, generator=gen, dtype=torch.float64) - 0.5)
    for n in range(num_cams):
        yaw = YAWS_DEG[n % len(YAWS_DEG)]
        K[n, 0, 0] = f
        K[n, 1, 1] = f
        K[n, 0, 2] = W / 2.0 + 16.0 * float(jit[n, 0])
        K[n, 1, 2] = H / 2.0 + 16.0 * float(jit[n, 1])
        K[n, 2, 2] = 1.0
        # small pitch/roll so that no rotation is axis aligned (exercises the FMA order)
        R = _rot("z", yaw) @ _rot("y", 3.0 * float(jit[n, 2])) @ _rot("x", 3.0 * float(jit[n, 3])) @ swap
        M[n, :3, :3] = R.to(torch.float32)
        y = math.radians(yaw)
        M[n, 3, 0] = 1.5 * math.cos(y)
        M[n, 3, 1] = 1.5 * math.sin(y)
        M[n, 3, 2] = -0.3
        M[n, 3, 3] = 1.0
    return K, M


def make_frame(frame_idx, H, W, num_cams=6, with_conf=True, scene="mixture"):
    """One synthetic frame.  Returns dict(depth, conf, sky, intrinsics, cam2lidar).

    scene="mixture" (SURVEY §8(d)): per pixel 85 % U(1,60) m, 10 % sky
    U(100.5,200) (removed by max_depth), 3 % exactly 0, 1 % NaN, 1 % +Inf.
    scene="ground": a ground plane 1.8 m below the cameras plus random walls,
    for realistic (clustered) voxel occupancy; same invalid-pixel mixture.
    """
    gen = torch.Generator().manual_seed(1234 + int(frame_idx))
    K, M = make_calibration(num_cams, H, W, gen)
    sel = torch.rand(num_cams, H, W, generator=gen)
    depth = 1.0 + 59.0 * torch.rand(num_cams, H, W, generator=gen)
    if scene == "ground":
        v = torch.arange(H, dtype=torch.float32).view(1, H, 1)
        cy = K[:, 1, 2].view(-1, 1, 1)
        fy = K[:, 1, 1].view(-1, 1, 1)
        ray = (v - cy) / fy                       # y_cam / z_cam (down positive)
        ground = torch.where(ray > 0.02, 1.8 / ray.clamp_min(0.02), torch.full_like(ray, 1e9))
        wall = 4.0 + 50.0 * torch.rand(num_cams, 1, W // 32 + 1, generator=gen)
        wall = wall.repeat_interleave(32, dim=2)[:, :, :W]
        depth = torch.minimum(ground.expand(num_cams, H, W), wall.expand(num_cams, H, W))
        depth = depth * (1.0 + 0.01 * (torch.rand(num_cams, H, W, generator=gen) - 0.5))
        depth = depth.contiguous()
    sky_d = 100.5 + 99.5 * torch.rand(num_cams, H, W, generator=gen)
    sky = sel < 0.10
    depth = torch.where(sky, sky_d, depth)
    depth = torch.where((sel >= 0.10) & (sel < 0.13), torch.zeros_like(depth), depth)
    depth = torch.where((sel >= 0.13) & (sel < 0.14), torch.full_like(depth, float("nan")), depth)
    depth = torch.where((sel >= 0.14) & (sel < 0.15), torch.full_like(depth, float("inf")), depth)
    out = dict(depth=depth.contiguous(), intrinsics=K, cam2lidar=M, sky=sky.contiguous())
    if with_conf:
        # DA3 head: conf = exp(x) + 1 >= 1 (depth_anything_3/model/dualdpt.py:349-354)
        e = torch.empty(num_cams, H, W).exponential_(1.0, generator=gen)
        out["conf"] = (1.0 + e).contiguous()
    return out


def make_batch(frame_ids, H, W, num_cams=6, with_conf=True, scene="mixture"):
    """Stack frames: depth (B,N,H,W), conf, sky (bool), intrinsics (B,N,3,3), cam2lidar (B,N,4,4)."""
    frames = [make_frame(i, H, W, num_cams, with_conf, scene) for i in frame_ids]
    return {k: torch.stack([f[k] for f in frames], dim=0) for k in frames[0]}
