"""Per-stage times (rd3_profile_*) of the fused path for any config / scene / batch: python tools/stage_probe.py C4 ground 64"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rd3_b200
from rd3_b200 import synthetic, _lib
cfgname, scene, B = sys.argv[1], sys.argv[2], int(sys.argv[3])
cfg = synthetic.CONFIGS[cfgname]
H, W = cfg["hw"]
b = synthetic.make_batch(list(range(B)), H, W, with_conf=False, scene=scene)
d = {k: v.cuda() for k, v in b.items()}
mod = rd3_b200.DepthToVoxels(cfg["voxel_size"], cfg["pcr"], cfg["max_points"], cfg["max_voxels"], max_depth=synthetic.MAX_DEPTH,
                             reuse_buffers=True).cuda().train()
for _ in range(3):
    r = mod(d["depth"], d["intrinsics"], d["cam2lidar"])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    mod(d["depth"], d["intrinsics"], d["cam2lidar"])
e1.record(); torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(5):
    mod(d["depth"], d["intrinsics"], d["cam2lidar"])
torch.cuda.synchronize()
ms, calls = _lib.profile_read()
_lib.profile_enable(False)
print(cfgname, scene, B, "ms/step %.3f" % (e0.elapsed_time(e1) / 20), {k: round(v / 5, 3) for k, v in ms.items()}, "voxels", r["voxel_num"][:4].tolist())
