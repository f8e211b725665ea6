"""CPU study of the fast cell decision's error bound (rd3_common.cuh: pixel_key_fast; depth.cu: calib_kernel).

For every valid pixel of a few synthetic frames: the reference cell coordinate f_ref (oracle unprojection, then
RN(RN(o - lo) / vs) in fp32) against the direct map f' = fma(z, fma(A, u, fma(B, v, C)), T) with the constants
calib_kernel derives, as a fraction of the analytic bound 2^-24 (z Qc + Pc).  The kernels decide a pixel only when
it is more than 2 x that bound (2^-23 ...) away from a cell boundary.  fp32 FMAs are emulated through fp64
(exact product, one extra rounding in 2^-29 of the cases: irrelevant for a maximum over ratios << 1).

    python tools/tol_study.py [frames] [H] [W]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from rd3_b200 import synthetic  # noqa: E402

nfr = int(sys.argv[1]) if len(sys.argv) > 1 else 2
H = int(sys.argv[2]) if len(sys.argv) > 2 else 504
W = int(sys.argv[3]) if len(sys.argv) > 3 else 896
vs = np.array([0.075, 0.075, 0.2], np.float32)
lo = np.array([-54.0, -54.0, -5.0], np.float32)
f32 = np.float32


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


worst = 0.0
hist = np.zeros(20, np.int64)
total = 0
for scene in ("mixture", "ground"):
    for fi in range(nfr):
        fr = synthetic.make_frame(900 + fi, H, W, scene=scene)
        depth, K, M = fr["depth"].numpy(), fr["intrinsics"].numpy(), fr["cam2lidar"].numpy()
        pts, pix = oracle.unproject(depth, K, M, max_depth=synthetic.MAX_DEPTH, return_pix=True)
        cam = pix // (H * W)
        rem = pix - cam * (H * W)
        v = (rem // W).astype(f32)
        u = (rem - (rem // W) * W).astype(f32)
        z = depth.reshape(-1)[pix]
        for a in range(3):
            f_ref = ((pts[:, a] - lo[a]).astype(f32) / vs[a]).astype(f32)
            ratio = np.zeros(len(pix))
            for c in range(depth.shape[0]):
                m = cam == c
                if not m.any():
                    continue
                fx, fy, cx, cy = (float(K[c, 0, 0]), float(K[c, 1, 1]), float(K[c, 0, 2]), float(K[c, 1, 2]))
                r0, r1, r2, t = (float(M[c, a, 0]), float(M[c, a, 1]), float(M[c, a, 2]), float(M[c, 3, a]))
                rv = 1.0 / float(vs[a])
                A, B = f32(r0 / fx * rv), f32(r1 / fy * rv)
                C = f32((r2 - r0 * cx / fx - r1 * cy / fy) * rv)
                T = f32((t - float(lo[a])) * rv - 0.5)
                ex = max(abs(cx), abs((W - 1) - cx)) / abs(fx)
                ey = max(abs(cy), abs((H - 1) - cy)) / abs(fy)
                D = abs(float(A)) * (W - 1) + abs(float(B)) * (H - 1) + abs(float(C))
                Q = (abs(r0) * ex + abs(r1) * ey + abs(r2)) * rv
                P = abs(float(T)) + 2.0 + 10.0 * abs(t) * rv + 3.0 * abs(float(lo[a])) * rv
                Qc, Pc = 3.0 * D + 10.0 * Q, P
                one = np.ones(int(m.sum()), f32)
                row = fma32(B * one, v[m], C * one)
                ax = fma32(A * one, u[m], row)
                h = fma32(z[m], ax, T * one)
                err = np.abs(h.astype(np.float64) - (f_ref[m].astype(np.float64) - 0.5))
                ratio[m] = err / (2.0 ** -24 * (z[m].astype(np.float64) * Qc + Pc))
            worst = max(worst, float(ratio.max()))
            hist += np.histogram(ratio, bins=20, range=(0.0, 1.0))[0]
            total += len(ratio)
print("pixels x axes: %d   max |f' - f_ref| / (2^-24 (z Qc + Pc)) = %.4f" % (total, worst))
print("share of samples per tenth of the bound:", np.round(hist.reshape(10, 2).sum(1) / total, 6).tolist())
