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


def to_sparse_encoder_inputs(result, batch_offset=0):
    """(voxel_features (sum M, F), coors (sum M, 4) [b,z,y,x], batch_size) of a (gathered)
    result: the cat + F.pad(coor, (1,0), value=i) of sparse_refinement.py:393-402 as one kernel
    (``rd3_pack_sparse_inputs``).  CUDA tensors only -- there is no CPU path."""
    from .fused import pack_sparse_inputs
    return pack_sparse_inputs(result, batch_offset=batch_offset)
