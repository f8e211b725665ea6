"""One DynamicScatter forward (C3: 2.3 M points on the 1440x1440x40 grid) for ncu launch lists."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rd3_b200
from rd3_b200 import synthetic
scene = sys.argv[1] if len(sys.argv) > 1 else "mixture"
c = synthetic.CONFIGS["C3"]
H, W = c["hw"]
b = synthetic.make_batch([0], H, W, scene=scene)
d = {k: v.cuda() for k, v in b.items()}
pts, cnt = rd3_b200.unproject_padded(d["depth"], d["intrinsics"], d["cam2lidar"], max_depth=synthetic.MAX_DEPTH)
cloud = pts[0, :int(cnt[0])].contiguous()
vox = rd3_b200.Voxelization(c["voxel_size"], c["pcr"], -1)
for red in (True, False):
    ds = rd3_b200.DynamicScatter(c["voxel_size"], c["pcr"], red)
    for _ in range(3):
        coors = vox(cloud)
        vf, vc = ds(cloud, coors)
torch.cuda.synchronize()
print(cloud.shape, vf.shape)
