"""Multi-GPU: frames (samples) are independent, so the path shards by sample with
no collective inside it (the reference loops samples in Python,
sparse_refinement.py:382-391, scatter_points.py:88-95).  One process per GPU;
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is used only to
gather the fixed-size padded outputs.
"""
import torch
import torch.distributed as dist


def shard_range(num_frames, world_size, rank):
    """Contiguous chunk of frames owned by ``rank``: sizes differ by at most one,
    earlier ranks get the larger chunks.  Returns (start, stop)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    base, rem = divmod(int(num_frames), world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(num_frames, world_size):
    return [shard_range(num_frames, world_size, r)[1] - shard_range(num_frames, world_size, r)[0]
            for r in range(world_size)]


def gather_voxel_outputs(result, num_frames_total=None, group=None):
    """all_gather the per-rank outputs of ``DepthToVoxels`` (padded, fixed size).

    ``result``: dict with voxels (b,MV,K,C), coors (b,MV,3), num_points (b,MV),
    voxel_mean (b,MV,F) | None, voxel_num (b,), where b is this rank's chunk of a
    ``num_frames_total`` batch sharded with :func:`shard_range`.
    Returns the same dict for the whole batch, frames in global order, on every rank.
    Ranks may own different numbers of frames (chunks are padded to the largest).
    """
    if not dist.is_available() or not dist.is_initialized():
        return result
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    b_local = result["voxel_num"].shape[0]
    if num_frames_total is None:
        counts = torch.tensor([b_local], device=result["voxel_num"].device)
        allc = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(allc, counts, group=group)
        sizes = [int(c.item()) for c in allc]
    else:
        sizes = shard_sizes(num_frames_total, world)
        if sizes[rank] != b_local:
            raise ValueError("rank %d holds %d frames, shard_range says %d" % (rank, b_local, sizes[rank]))
    bmax = max(sizes)
    out = {}
    for name, t in result.items():
        if t is None:
            out[name] = None
            continue
        if t.shape[0] < bmax:                                     # pad the short chunks
            pad = t.new_zeros((bmax - t.shape[0],) + tuple(t.shape[1:]))
            t = torch.cat([t, pad], dim=0)
        buf = t.new_empty((world * bmax,) + tuple(t.shape[1:]))      # concatenated along dim 0
        dist.all_gather_into_tensor(buf, t.contiguous(), group=group)
        buf = buf.view((world, bmax) + tuple(t.shape[1:]))
        out[name] = torch.cat([buf[r, :sizes[r]] for r in range(world)], dim=0)
    return out


def flat_layout(num_frames, max_voxels, feat_dim=3):
    """Element offsets of the regions of ``DepthToVoxels(flat_outputs=True)``'s int32 buffer for ``num_frames`` frames:
    dict name -> (start, stop, shape, dtype), plus ``total`` under the key ``None``."""
    n = int(num_frames) * int(max_voxels)
    return {"voxel_mean": (0, n * feat_dim, (num_frames, max_voxels, feat_dim), torch.float32),
            "coors": (n * feat_dim, n * feat_dim + n * 3, (num_frames, max_voxels, 3), torch.int32),
            "num_points": (n * feat_dim + n * 3, n * feat_dim + n * 4, (num_frames, max_voxels), torch.int32),
            "voxel_num": (n * feat_dim + n * 4, n * feat_dim + n * 4 + num_frames, (num_frames,), torch.int32),
            None: n * feat_dim + n * 4 + num_frames}


def gather_flat_outputs(result, max_voxels, out=None, group=None, async_op=False):
    """ONE collective for what the sparse encoder consumes of every rank's shard: all-gathers ``result["flat"]``
    (``DepthToVoxels(flat_outputs=True)``: mean | coors | num | voxel_num of this rank's frames; equal shards on all
    ranks) into ``out`` (world, flat) -- allocated when None -- and returns ``(views, work)``: ``views`` is the same dict
    as ``gather_voxel_outputs`` gives (frames in global order, no copy: they alias ``out``), ``work`` the async handle
    or None.  bench.py times this pattern with two buffer sets on a communication stream."""
    flat = result["flat"]
    b_local = result["voxel_num"].shape[0]
    lay = flat_layout(b_local, max_voxels, result["voxel_mean"].shape[-1])
    if flat.numel() != lay[None]:
        raise ValueError("result['flat'] has %d elements, the layout for %d frames x %d voxels has %d"
                         % (flat.numel(), b_local, max_voxels, lay[None]))
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if out is None:
        out = flat.new_empty((world, flat.numel()))
    work = None
    if world == 1:
        out[0].copy_(flat)
    else:
        work = dist.all_gather_into_tensor(out.view(-1), flat, group=group, async_op=async_op)
    views = {}
    for name in ("voxel_mean", "coors", "num_points", "voxel_num"):
        a, b, shape, dtype = lay[name]
        v = out[:, a:b]
        if dtype != torch.int32:
            v = v.view(dtype)
        views[name] = v.reshape((world * shape[0],) + tuple(shape[1:]))     # rank-major == global frame order
    views["voxels"] = None
    return views, (work if async_op else None)


def to_sparse_encoder_inputs(result, batch_offset=0):
    """(voxel_features (sum M, F), coors (sum M, 4) [b,z,y,x], batch_size) of a (gathered)
    result: the cat + F.pad(coor, (1,0), value=i) of sparse_refinement.py:393-402 as one kernel
    (``rd3_pack_sparse_inputs``).  CUDA tensors only -- there is no CPU path."""
    from .fused import pack_sparse_inputs
    return pack_sparse_inputs(result, batch_offset=batch_offset)
