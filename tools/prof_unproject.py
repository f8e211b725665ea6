"""One unprojection call at C2 shape (8 frames) -- the command profiled under ncu for profiles/r2_unproject_full.txt."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rd3_b200
from rd3_b200 import synthetic
H, W = synthetic.CONFIGS["C2"]["hw"]
b = synthetic.make_batch(list(range(8)), H, W)
d = {k: v.cuda() for k, v in b.items()}
for _ in range(3):
    pts, counts = rd3_b200.unproject_padded(d["depth"], d["intrinsics"], d["cam2lidar"], max_depth=synthetic.MAX_DEPTH)
torch.cuda.synchronize()
print(counts.tolist())
