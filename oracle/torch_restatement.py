"""torch-CPU restatement of the Python-level reference code on the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference modules themselves cannot be imported here: ``mmdet3d`` hard-
imports mmcv/mmdet/mmseg (mmdetection3d/mmdet3d/__init__.py:2-5), none of
which are installed.  Each function below performs the same torch operations,
in the same order and dtype, as the cited reference lines, so that on CPU
tensors it produces the bits the reference would.
"""
import torch


def backproject_depth_to_points(depths, intrinsics, cam2lidar_rts,
                                max_depth=None, confs=None, conf_thresh=None,
                                sky_masks=None):
    """reconstruction_backbone.py:285-386 without the colour branch.

    depths (B,N,H,W), intrinsics (B,N,3,3), cam2lidar_rts (B,N,4,4).
    Optional conf / sky masks follow the commented block at :341-346 whose live
    form is tools/inference_nuscenes.py:399-414.
    Returns list_B of (P_b, 3) tensors.
    """
    B, N, H, W = depths.shape
    dt = depths.dtype
    u = torch.arange(W, dtype=dt)
    v = torch.arange(H, dtype=dt)
    vv, uu = torch.meshgrid(v, u, indexing="ij")                       # :312-314
    out = []
    for b in range(B):
        per_cam = []
        for n in range(N):
            z = depths[b, n]
            K = intrinsics[b, n]
            fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]        # :325-326
            x = (uu - cx) * z / fx                                      # :329-331
            y = (vv - cy) * z / fy
            pts = torch.stack([x.reshape(-1), y.reshape(-1), z.reshape(-1)], dim=1)
            zf = z.reshape(-1)
            valid = (zf > 0) & torch.isfinite(zf)                       # :338
            if max_depth is not None:
                valid = valid & (zf <= max_depth)                       # :339-340
            if confs is not None and conf_thresh is not None:
                valid = valid & (confs[b, n].reshape(-1) >= conf_thresh)
            if sky_masks is not None:
                valid = valid & (~sky_masks[b, n].reshape(-1))
            pts = pts[valid]                                            # :348
            if pts.numel() > 0:
                M = cam2lidar_rts[b, n]
                pts = pts @ M[:3, :3].T + M[3, :3]                      # :370
                per_cam.append(pts)
        out.append(torch.cat(per_cam, dim=0) if per_cam
                   else torch.zeros((0, 3), dtype=dt))                  # :376-381
    return out


def filter_point_by_range(points, point_cloud_range):
    """respoint_post_processing.py:190-199: inclusive on both ends."""
    x0, y0, z0, x1, y1, z1 = point_cloud_range
    m = ((points[:, 0] >= x0) & (points[:, 0] <= x1) &
         (points[:, 1] >= y0) & (points[:, 1] <= y1) &
         (points[:, 2] >= z0) & (points[:, 2] <= z1))
    return points[m], torch.nonzero(m, as_tuple=False).squeeze(1)


def extract_head_outputs(depth, conf=None, sky=None):
    """depth_anything_3/utils/io/output_processor.py:79-101,152-168: the heads emit (B, N, H, W, 1);
    depth / conf are squeezed, the sky probability becomes the boolean mask ``sky >= 0.5``."""
    d = depth.squeeze(-1)
    c = conf.squeeze(-1) if conf is not None else None
    s = (sky.squeeze(-1) >= 0.5) if sky is not None else None
    return d, c, s


def conf_threshold(conf, sky, percentile):
    """tools/inference_nuscenes.py:351-361 (numpy percentile, linear interp).

    conf (N,H,W) float, sky (N,H,W) bool or None."""
    import numpy as np
    c = conf.numpy() if torch.is_tensor(conf) else np.asarray(conf)
    if sky is not None:
        s = sky.numpy() if torch.is_tensor(sky) else np.asarray(sky)
        px = c[~s] if (~s).sum() > 10 else c.flatten()
    else:
        px = c.flatten()
    return float(np.percentile(px, percentile))


def hard_simple_vfe(features, num_points, num_features):
    """voxel_encoder.py:45-46."""
    mean = features[:, :, :num_features].sum(dim=1, keepdim=False) / \
        num_points.type_as(features).view(-1, 1)
    return mean.contiguous()


def voxelization_forward(ref_layer, points, voxel_size, coors_range,
                         max_points, max_voxels):
    """voxelize.py:52-70 driving the compiled reference op (oracle/_ref)."""
    if max_points == -1 or max_voxels == -1:
        coors = points.new_zeros(size=(points.size(0), 3), dtype=torch.int)
        ref_layer.dynamic_voxelize(points, coors, voxel_size, coors_range, 3)
        return coors
    voxels = points.new_zeros(size=(max_voxels, max_points, points.size(1)))
    coors = points.new_zeros(size=(max_voxels, 3), dtype=torch.int)
    num = points.new_zeros(size=(max_voxels,), dtype=torch.int)
    n = ref_layer.hard_voxelize(points, voxels, coors, num, voxel_size,
                                coors_range, max_points, max_voxels, 3, True)
    return voxels[:n], coors[:n], num[:n]


def dynamic_point_to_voxel_forward(feats, coors, reduce_type):
    """scatter_points_cuda.cu:183-239 with torch-CPU ops (faithful unique_dim).

    Returns [voxel_feats, voxel_coors, point2voxel_map int32, count int32].
    Sums run through index_add_ in fp32 (point order on one thread)."""
    if reduce_type not in ("sum", "mean", "max"):
        raise RuntimeError("do not support reduce type " + reduce_type)
    N, C = feats.shape
    if N == 0:                                                          # :192-196
        return [feats.clone().detach(), coors.clone().detach(),
                coors.new_empty((0,), dtype=torch.int32),
                coors.new_empty((0,), dtype=torch.int32)]
    clean = coors.masked_fill(coors.lt(0).any(-1, True), -1)            # :202
    out_coors, cmap, cnt = torch.unique(clean, dim=0, sorted=True,
                                        return_inverse=True,
                                        return_counts=True)             # :204-205
    if bool(out_coors[0, 0].lt(0)):                                     # :207-212
        out_coors = out_coors[1:]
        cnt = cnt[1:]
        cmap = cmap - 1
    cmap = cmap.to(torch.int32)
    cnt = cnt.to(torch.int32)
    M = out_coors.size(0)
    keep = cmap >= 0
    idx = cmap[keep].long()
    if reduce_type == "max":
        red = feats.new_full((M, C), float("-inf"))
        red.scatter_reduce_(0, idx.view(-1, 1).expand(-1, C), feats[keep],
                            reduce="amax", include_self=True)
    else:
        red = feats.new_zeros((M, C))
        red.index_add_(0, idx, feats[keep])
        if reduce_type == "mean":
            red /= cnt.unsqueeze(-1).to(red.dtype)                      # :233-234
    return [red, out_coors, cmap, cnt]


def dynamic_point_to_voxel_backward(grad_voxel, feats, voxel_feats, cmap, cnt,
                                    reduce_type):
    """scatter_points_cuda.cu:105-179,241-308 -> grad_feats (N,C)."""
    N, C = feats.shape
    g = torch.zeros_like(feats)                                         # :259
    M = voxel_feats.size(0)
    if N == 0 or M == 0:
        return g
    keep = cmap >= 0
    idx = cmap[keep].long()
    if reduce_type == "sum":
        g[keep] = grad_voxel[idx]
    elif reduce_type == "mean":
        g[keep] = grad_voxel[idx] / cnt[idx].to(g.dtype).unsqueeze(-1)
    else:
        # lowest point index attaining the max per (voxel, feature) (:149-152)
        pts = torch.nonzero(keep).squeeze(1)
        hit = feats[pts] == voxel_feats[idx]
        src = torch.full((M, C), N, dtype=torch.long)
        pidx = pts.view(-1, 1).expand(-1, C)
        cand = torch.where(hit, pidx, torch.full_like(pidx, N))
        src.scatter_reduce_(0, idx.view(-1, 1).expand(-1, C), cand,
                            reduce="amin", include_self=True)
        cols = torch.arange(C).view(1, -1).expand(M, -1)
        ok = src < N
        g[src[ok], cols[ok]] = grad_voxel[ok]
    return g


def dynamic_scatter_batched(feats, coors, reduce_type):
    """scatter_points.py:78-99: (N,4) [b,z,y,x] coors -> per-batch scatter, cat."""
    if coors.size(-1) == 3:
        r = dynamic_point_to_voxel_forward(feats.contiguous(),
                                           coors.contiguous(), reduce_type)
        return r[0], r[1]
    batch_size = int(coors[-1, 0]) + 1
    vs, cs = [], []
    for i in range(batch_size):
        inds = torch.where(coors[:, 0] == i)
        r = dynamic_point_to_voxel_forward(feats[inds].contiguous(),
                                           coors[inds][:, 1:].contiguous(),
                                           reduce_type)
        cs.append(torch.nn.functional.pad(r[1], (1, 0), mode="constant", value=i))
        vs.append(r[0])
    return torch.cat(vs, dim=0), torch.cat(cs, dim=0)


def voxel_downsample(ref_layer, points, voxel_size, point_cloud_range, colors=None):
    """respoint_post_processing.py:29-98 on CPU tensors, driving the compiled reference op."""
    pcr = point_cloud_range
    if pcr is None:
        pcr = (points.min(dim=0).values - 1.0).tolist() + (points.max(dim=0).values + 1.0).tolist()
    vs = [voxel_size] * 3 if isinstance(voxel_size, (int, float)) else list(voxel_size)
    voxels, coors, num = voxelization_forward(ref_layer, points.contiguous(), vs, list(pcr), 100, 200000)
    if voxels.shape[0] == 0:
        return points, colors, torch.arange(points.shape[0])
    centers = torch.stack([voxels[i, :int(num[i])].mean(dim=0) for i in range(voxels.shape[0])], dim=0)
    idx = torch.arange(centers.shape[0])
    vcol = None
    if colors is not None:
        idx = torch.argmin(torch.cdist(centers, points, compute_mode='donot_use_mm_for_euclid_dist'), dim=1)
        vcol = colors[idx]
    return centers, vcol, idx


def get_paddings_indicator(actual_num, max_num, axis=0):
    """mmdet3d/models/voxel_encoders/utils.py:9-29."""
    actual_num = torch.unsqueeze(actual_num, axis + 1)
    max_num_shape = [1] * len(actual_num.shape)
    max_num_shape[axis + 1] = -1
    max_num = torch.arange(max_num, dtype=torch.int, device=actual_num.device).view(max_num_shape)
    return actual_num.int() > max_num


def pillar_feature_decorations(features, num_points, coors, voxel_size=(0.2, 0.2, 4),
                               point_cloud_range=(0, -40, -3, 70.4, 40, 1), with_cluster_center=True,
                               with_voxel_center=True, with_distance=False, legacy=True):
    """PillarFeatureNet.__init__/forward up to the PFN layers, pillar_encoder.py:84-143,
    statement for statement (the legacy branch aliases ``features`` exactly like the reference)."""
    vx, vy = voxel_size[0], voxel_size[1]
    x_offset = vx / 2 + point_cloud_range[0]                            # :87-88
    y_offset = vy / 2 + point_cloud_range[1]
    features = features.clone()
    features_ls = [features]
    if with_cluster_center:                                             # :108-113
        points_mean = features[:, :, :3].sum(dim=1, keepdim=True) / num_points.type_as(features).view(-1, 1, 1)
        f_cluster = features[:, :, :3] - points_mean
        features_ls.append(f_cluster)
    dtype = features.dtype
    if with_voxel_center:                                               # :117-134
        if not legacy:
            f_center = torch.zeros_like(features[:, :, :2])
            f_center[:, :, 0] = features[:, :, 0] - (coors[:, 3].to(dtype).unsqueeze(1) * vx + x_offset)
            f_center[:, :, 1] = features[:, :, 1] - (coors[:, 2].to(dtype).unsqueeze(1) * vy + y_offset)
        else:
            f_center = features[:, :, :2]
            f_center[:, :, 0] = f_center[:, :, 0] - (coors[:, 3].type_as(features).unsqueeze(1) * vx + x_offset)
            f_center[:, :, 1] = f_center[:, :, 1] - (coors[:, 2].type_as(features).unsqueeze(1) * vy + y_offset)
        features_ls.append(f_center)
    if with_distance:                                                   # :136-138
        points_dist = torch.norm(features[:, :, :3], 2, 2, keepdim=True)
        features_ls.append(points_dist)
    features = torch.cat(features_ls, dim=-1)                           # :141
    voxel_count = features.shape[1]
    mask = get_paddings_indicator(num_points, voxel_count, axis=0)      # :146-148
    mask = torch.unsqueeze(mask, -1).type_as(features)
    features *= mask
    return features


def point_pillars_scatter(voxel_features, coors, in_channels, ny, nx, batch_size=None):
    """PointPillarsScatter.forward, pillar_scatter.py:27-102."""
    if batch_size is None:                                              # forward_single :39-60
        canvas = torch.zeros(in_channels, nx * ny, dtype=voxel_features.dtype)
        indices = (coors[:, 1] * nx + coors[:, 2]).long()
        canvas[:, indices] = voxel_features.t()
        return [canvas.view(1, in_channels, ny, nx)]
    batch_canvas = []                                                   # forward_batch :62-102
    for batch_itt in range(batch_size):
        canvas = torch.zeros(in_channels, nx * ny, dtype=voxel_features.dtype)
        batch_mask = coors[:, 0] == batch_itt
        this_coors = coors[batch_mask, :]
        indices = (this_coors[:, 2] * nx + this_coors[:, 3]).type(torch.long)
        canvas[:, indices] = voxel_features[batch_mask, :].t()
        batch_canvas.append(canvas)
    return torch.stack(batch_canvas, 0).view(batch_size, in_channels, ny, nx)


def map_voxel_center_to_point(pts_coors, voxel_mean, voxel_coors, voxel_size, point_cloud_range):
    """DynamicVFE.map_voxel_center_to_point, voxel_encoder.py:179-219 (dense canvas and all)."""
    vx, vy, vz = voxel_size
    canvas_z = int((point_cloud_range[5] - point_cloud_range[2]) / vz)          # :192-197
    canvas_y = int((point_cloud_range[4] - point_cloud_range[1]) / vy)
    canvas_x = int((point_cloud_range[3] - point_cloud_range[0]) / vx)
    batch_size = pts_coors[-1, 0] + 1
    canvas_len = canvas_z * canvas_y * canvas_x * batch_size
    canvas = voxel_mean.new_zeros(canvas_len, dtype=torch.long)
    indices = (voxel_coors[:, 0] * canvas_z * canvas_y * canvas_x + voxel_coors[:, 1] * canvas_y * canvas_x +
               voxel_coors[:, 2] * canvas_x + voxel_coors[:, 3])
    canvas[indices.long()] = torch.arange(start=0, end=voxel_mean.size(0))
    voxel_index = (pts_coors[:, 0] * canvas_z * canvas_y * canvas_x + pts_coors[:, 1] * canvas_y * canvas_x +
                   pts_coors[:, 2] * canvas_x + pts_coors[:, 3])
    voxel_inds = canvas[voxel_index.long()]
    return voxel_mean[voxel_inds, ...]


def conf_threshold_numpy1(conf, sky, percentile):
    """The same selection as :func:`conf_threshold` with the arithmetic NumPy < 2 (the reference's
    pin, requirements.txt:8) performs inside ``np.percentile`` (numpy/lib/function_base.py,
    method 'linear'): q/100 and the virtual index (n-1)*q in fp64, the difference of the two
    fp32 order statistics in fp32, the interpolation in fp64.  Pinned by source only: the
    installed numpy is 2.x, whose result :func:`conf_threshold` returns."""
    import numpy as np
    c = conf.numpy() if torch.is_tensor(conf) else np.asarray(conf)
    if sky is not None:
        s = sky.numpy() if torch.is_tensor(sky) else np.asarray(sky)
        px = c[~s] if (~s).sum() > 10 else c.flatten()
    else:
        px = c.flatten()
    n = px.size
    if n == 0:
        return float("nan")
    vidx = np.float64(n - 1) * (np.float64(percentile) / np.float64(100))
    prev, nxt = np.floor(vidx), np.floor(vidx) + 1
    gamma = vidx - prev
    if vidx >= n - 1:
        prev = nxt = n - 1
    if vidx < 0:
        prev = nxt = 0
    srt = np.sort(px.astype(np.float32))
    a, b = srt[int(prev)], srt[int(nxt)]
    d = np.float64(np.float32(b) - np.float32(a))
    r = np.float64(a) + d * gamma
    if gamma >= 0.5:
        r = np.float64(b) - d * (1 - gamma)
    return float(r)


def soft_voxel_occupancy(features, num_points, lambda_n=0.3, gamma_var=5.0, eps=1e-6):
    """SoftVoxelOccupancyVFE.forward, voxel_occupancy_encoder.py:60-99."""
    N, M, C = features.shape
    xyz = features[:, :, :3]
    mask = (torch.arange(M).unsqueeze(0) < num_points.unsqueeze(1))
    mask_exp = mask.unsqueeze(-1).float()
    xyz_sum = (xyz * mask_exp).sum(dim=1)
    denom = num_points.unsqueeze(1).float() + eps
    xyz_mean = xyz_sum / denom
    diff = (xyz - xyz_mean.unsqueeze(1)) * mask_exp
    var = (diff.pow(2).sum(dim=1) / denom).mean(dim=1)
    n = num_points.float()
    occupancy = 1.0 - torch.exp(-lambda_n * n - gamma_var * var)
    return occupancy.view(-1, 1).contiguous()


def occupancy_feature_map(voxel_occupancy, coors_list, B, Z, Y, X):
    """sparse_refinement.py:566-587: dense (B, Z, Y, X) map from per-sample (z,y,x) coors."""
    per_batch = torch.split(voxel_occupancy, [c.shape[0] for c in coors_list], dim=0)
    out = torch.zeros(B, Z, Y, X)
    for b_idx in range(B):
        c = coors_list[b_idx]
        out[b_idx, c[:, 0].long(), c[:, 1].long(), c[:, 2].long()] = per_batch[b_idx].squeeze(-1)
    return out
