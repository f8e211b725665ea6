"""Full-size soak: fused depth->voxels of many C2 frames (2.7 M pixels each, both scenes) against
the C oracle, bit for bit.     python tools/soak_full.py [first_frame] [count]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import rd3_b200  # noqa: E402
from rd3_b200 import synthetic  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 16
c = synthetic.CONFIGS["C2"]
H, W = c["hw"]
dev = "cuda:0"
mod = rd3_b200.DepthToVoxels(c["voxel_size"], c["pcr"], c["max_points"], c["max_voxels"],
                             max_depth=synthetic.MAX_DEPTH).to(dev).train()
t0 = time.time()
bad = 0
for k in range(0, count, 4):
    ids = list(range(first + k, first + min(k + 4, count)))
    for scene in ("mixture", "ground"):
        b = synthetic.make_batch(ids, H, W, with_conf=False, scene=scene)
        r = mod(b["depth"].to(dev), b["intrinsics"].to(dev), b["cam2lidar"].to(dev))
        vn = r["voxel_num"].cpu().numpy()
        for i, fid in enumerate(ids):
            pts = oracle.unproject(b["depth"][i].numpy(), b["intrinsics"][i].numpy(), b["cam2lidar"][i].numpy(),
                                   max_depth=synthetic.MAX_DEPTH)
            ov, oc, on = oracle.hard_voxelize(pts, list(c["voxel_size"]), list(c["pcr"]), c["max_points"],
                                              c["max_voxels"][0])
            m = int(vn[i])
            ok = (m == len(oc) and np.array_equal(r["coors"][i, :m].cpu().numpy(), oc) and
                  np.array_equal(r["num_points"][i, :m].cpu().numpy(), on) and
                  np.array_equal(r["voxels"][i, :m].cpu().numpy().view(np.uint32), ov.view(np.uint32)))
            if not ok:
                bad += 1
                print("MISMATCH frame", fid, scene, m, len(oc))
print("%d frames x 2 scenes x %d pixels: %d mismatching frames in %.0f s" % (count, 6 * H * W, bad, time.time() - t0))
sys.exit(1 if bad else 0)
