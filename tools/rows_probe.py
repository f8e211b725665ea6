"""One call each of the point-array operators at their BASELINE configs (C3 DynamicScatter, C2 / C4 hard
voxelization of a stored cloud, unprojection, conf percentile): the command whose kernel launch list is captured with
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rows_launches.csv python tools/rows_probe.py
(durations under ncu are cold-cache and serialised: they show which launches a row's time is made of)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rd3_b200  # noqa: E402
from rd3_b200 import synthetic, voxel_layer  # noqa: E402

dev = torch.device("cuda:0")
c2, c4 = synthetic.CONFIGS["C2"], synthetic.CONFIGS["C4"]
H, W = c2["hw"]
b = synthetic.make_batch([0], H, W)
d = {k: v.to(dev) for k, v in b.items()}
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(reps):
    torch.cuda.nvtx.range_push("unproject")
    pts, counts = rd3_b200.unproject_padded(d["depth"], d["intrinsics"], d["cam2lidar"], max_depth=synthetic.MAX_DEPTH)
    torch.cuda.nvtx.range_pop()
    cloud = pts[0, :int(counts[0])].contiguous()
    coors = rd3_b200.Voxelization(c2["voxel_size"], c2["pcr"], -1)(cloud)
    for avg in (True, False):
        rd3_b200.DynamicScatter(c2["voxel_size"], c2["pcr"], avg)(cloud, coors)
    for cfg in (c2, c4):
        K, mv = cfg["max_points"], cfg["max_voxels"][0]
        voxels = torch.empty((mv, K, 3), device=dev)
        co = torch.empty((mv, 3), dtype=torch.int32, device=dev)
        nu = torch.empty((mv,), dtype=torch.int32, device=dev)
        mean = torch.empty((mv, 3), device=dev)
        voxel_layer.hard_voxelize(cloud, voxels, co, nu, list(cfg["voxel_size"]), list(cfg["pcr"]), K, mv, voxel_mean=mean)
    rd3_b200.conf_threshold(d["conf"], d["sky"], 30.0)
    torch.cuda.synchronize()
print("ok")
