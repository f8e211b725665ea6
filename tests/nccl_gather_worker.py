"""Worker of tests/test_gpu_parity.py::test_nccl_gather_matches_single_process (launched with torch.distributed.run,
one rank per GPU): every rank voxelizes its shard of a small batch, the padded outputs are all-gathered over NCCL
(rd3_b200.parallel.gather_voxel_outputs) and every rank compares them, bit for bit, with its own single-GPU result
for the whole batch."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rd3_b200  # noqa: E402
from rd3_b200 import parallel, synthetic  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    c = synthetic.CONFIGS["C1"]
    BT, H, W = 5, 56, 96                                    # 5 frames over 2 ranks: unequal shards (3 + 2)
    host = synthetic.make_batch(list(range(BT)), H, W, with_conf=False)
    full = {k: v.to(dev) for k, v in host.items()}
    mod = rd3_b200.DepthToVoxels(c["voxel_size"], c["pcr"], c["max_points"], 3000, max_depth=synthetic.MAX_DEPTH).to(dev)
    ref = mod(full["depth"], full["intrinsics"], full["cam2lidar"])
    f0, f1 = parallel.shard_range(BT, world, rank)
    mine = mod(full["depth"][f0:f1].contiguous(), full["intrinsics"][f0:f1].contiguous(), full["cam2lidar"][f0:f1].contiguous())
    got = parallel.gather_voxel_outputs(mine, num_frames_total=BT)
    torch.cuda.synchronize()
    assert torch.equal(got["voxel_num"], ref["voxel_num"]), (got["voxel_num"], ref["voxel_num"])
    for i in range(BT):
        m = int(ref["voxel_num"][i])
        assert m > 100
        for k in ("voxels", "coors", "num_points", "voxel_mean"):
            assert torch.equal(got[k][i, :m], ref[k][i, :m]), (rank, i, k)
    feats, coors4, bs = parallel.to_sparse_encoder_inputs(got)
    assert bs == BT and coors4.shape[0] == int(ref["voxel_num"].sum()) and int(coors4[-1, 0]) == BT - 1
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok" % rank)


if __name__ == "__main__":
    main()
