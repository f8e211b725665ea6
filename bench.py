#!/usr/bin/env python
"""bench.py -- fused depth unprojection + hard voxelization (+ voxel mean) throughput.

Workload (BASELINE.json configs[1], "C2"): per GPU a batch of synthetic nuScenes-
shaped frames, 6 cameras x 504 x 896 DA3 depth (2 709 504 pixels/frame), voxel
(0.075, 0.075, 0.2) m, range (-54,-54,-5,54,54,3), max_points 10, max_voxels
120 000, max_depth 100.  One "step" = one pass of the hot path
over the whole batch.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's sm_100a path
    python bench.py --impl reference [--gpus N --steps K --warmup W] # reference CPU path, host cores

Prints ONE JSON line (rank 0).  `value` is pixels(points)/s of the whole job with the
inputs resident in HBM: at N > 1 the step's 64 frames are sharded by sample (BASELINE
config C5, strong scaling) and the NCCL all-gather of what the sparse encoder consumes is
inside the timed region.  `e2e` is the same path through the public API from pinned HOST
buffers with every host<->device copy inside the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "points_per_sec_unproject_voxelize_scatter"
UNIT = "points/s"
WORKLOAD = "C2"
WORKLOAD_DESC = ("C2: fused depth unprojection + hard voxelization + voxel mean, 6x504x896 depth, "
                 "voxel 0.075/0.075/0.2, max_points 10, max_voxels 120000, max_depth 100")


# --------------------------------------------------------------------------------------
# reference / CPU baseline (the ONLY place bench.py touches oracle/)
# --------------------------------------------------------------------------------------
def _cpu_worker_init(hw, cfg_name, frame_ids, nthreads):
    global _W
    import torch
    torch.set_num_threads(nthreads)
    import oracle
    from oracle import torch_restatement as tr
    from rd3_b200 import synthetic
    ref = oracle.ref_voxel_layer()
    frames = [synthetic.make_frame(i, hw[0], hw[1], with_conf=False) for i in frame_ids]
    _W = dict(torch=torch, oracle=oracle, tr=tr, ref=ref, frames=frames,
              cfg=synthetic.CONFIGS[cfg_name], max_depth=synthetic.MAX_DEPTH)


def _cpu_one_frame(f):
    """The reference's CPU path for one frame: torch-CPU unprojection
    (reconstruction_backbone.py:305-386) -> hard_voxelize_cpu (the reference's own C++,
    oracle/_ref; or the C port when that was never built) -> HardSimpleVFE."""
    W = _W
    torch, tr, cfg = W["torch"], W["tr"], W["cfg"]
    pts = tr.backproject_depth_to_points(f["depth"][None], f["intrinsics"][None], f["cam2lidar"][None],
                                         max_depth=W["max_depth"])[0]
    mv = cfg["max_voxels"][0]
    if W["ref"] is not None:
        v, c, n = tr.voxelization_forward(W["ref"], pts.contiguous(), list(cfg["voxel_size"]),
                                          list(cfg["pcr"]), cfg["max_points"], mv)
    else:
        v, c, n = W["oracle"].hard_voxelize(pts.numpy(), list(cfg["voxel_size"]), list(cfg["pcr"]),
                                            cfg["max_points"], mv)
        v, n = torch.from_numpy(v), torch.from_numpy(n)
    mean = tr.hard_simple_vfe(v, n, 3)
    return int(mean.shape[0])


def _cpu_worker_run(_):
    t0 = time.perf_counter()
    m = [_cpu_one_frame(f) for f in _W["frames"]]
    return time.perf_counter() - t0, m


class CpuReference:
    """Process-parallel reference path: one frame per worker per step (hard_voxelize_cpu is
    single-threaded by construction, so the host is filled with independent frames)."""

    def __init__(self, cfg_name, workers=None, frames_per_worker=1):
        import multiprocessing as mp
        from rd3_b200 import synthetic
        import oracle
        ncpu = os.cpu_count() or 1
        self.workers = workers or max(1, min(ncpu, 32))
        self.frames_per_worker = frames_per_worker
        self.kind = "reference" if oracle.ref_voxel_layer() is not None else "port"
        hw = synthetic.CONFIGS[cfg_name]["hw"]
        self.pix_per_frame = 6 * hw[0] * hw[1]
        ctx = mp.get_context("spawn")
        self.pools = []
        for w in range(self.workers):
            ids = [10000 + w * frames_per_worker + k for k in range(frames_per_worker)]
            self.pools.append(ctx.Pool(1, initializer=_cpu_worker_init, initargs=(hw, cfg_name, ids, 1)))
        # make sure every worker finished its init before anything is timed
        for p in self.pools:
            p.apply(int, (0,))

    def step(self):
        """Run one bounded sample on all workers; returns (wall seconds, frames)."""
        t0 = time.perf_counter()
        res = [p.apply_async(_cpu_worker_run, (0,)) for p in self.pools]
        out = [r.get() for r in res]
        wall = time.perf_counter() - t0
        return wall, self.workers * self.frames_per_worker, out

    def close(self):
        for p in self.pools:
            p.terminate()

    def describe(self):
        return "%d frames (one 6x504x896 frame per worker process, %d workers): torch-CPU unprojection + %s hard_voxelize_cpu + HardSimpleVFE" % (
            self.workers * self.frames_per_worker, self.workers,
            "the reference's compiled" if self.kind == "reference" else "C-port")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ref = CpuReference(WORKLOAD, workers=args.cpu_workers)
    for _ in range(args.warmup):
        ref.step()
    t = 0.0
    frames = 0
    for _ in range(args.steps):
        w, f, _o = ref.step()
        t += w
        frames += f
    ref.close()
    value = frames * ref.pix_per_frame / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "frames_per_sec": frames / t,
        "config": {"workload": WORKLOAD_DESC, "scene": "mixture", "frames_per_step": ref.workers,
                   "note": "bounded sample of the same workload: one frame per host core per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.workers, "kind": ref.kind,
                         "sample": ref.describe()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)
    return 0


# --------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------
_REAL_STDOUT = None


def _claim_stdout():
    """Everything libraries print (e.g. NCCL's version banner) goes to stderr; the one JSON
    line is written to the real stdout by _emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def _source_hash():
    """Stamp of the build that numbers measured offline (ncu traffic) belong to: the md5 of the machine code that
    build.py writes next to the library (unchanged by comment-only edits); a hash of the CUDA sources if it is missing."""
    import hashlib
    f = os.path.join(ROOT, "3d-reconstruction-detection_b200", "librd3_b200.sass_md5")
    if os.path.exists(f):
        return "sass:" + open(f).read().strip()
    h = hashlib.sha256()
    d = os.path.join(ROOT, "3d-reconstruction-detection_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    return "src:" + h.hexdigest()[:16]


def _shutdown_process_group(dist):
    """Leave the job without hanging.  The step was replayed as a CUDA graph that holds NCCL work; tearing the
    communicator down under it (dist.destroy_process_group, or the graph's destructor at interpreter exit) was seen to
    block until the launcher's timeout.  All ranks meet at a barrier -- the line is printed by then -- and the process
    ends with exit code 0 without running the teardown; a watchdog does the same if the barrier itself blocks."""
    import torch

    def _bail():
        time.sleep(30.0)
        os._exit(0)

    threading.Thread(target=_bail, daemon=True).start()
    torch.cuda.synchronize()
    try:
        dist.barrier()
        torch.cuda.synchronize()
    except Exception:                                          # noqa: BLE001
        pass
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=64, help="frames per step in TOTAL (BASELINE config C5: sharded over the GPUs)")
    ap.add_argument("--cpu-workers", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-rows", action="store_true", help="skip the C1 / C3 / C4 operator rows")
    ap.add_argument("--no-masks", action="store_true", help="skip the conf + sky + percentile variant")
    ap.add_argument("--no-graph", action="store_true", help="N > 1: do not replay the step as a CUDA graph")
    ap.add_argument("--scene", default="mixture", choices=["mixture", "ground"],
                    help="synthetic depth: SURVEY 8(d) per-pixel mixture (default) or a structured ground+walls scene")
    ap.add_argument("--e2e-chunk", type=int, default=8, help="frames per H2D -> kernels -> D2H chunk of the e2e leg")
    ap.add_argument("--e2e-streams", type=int, default=3)
    ap.add_argument("--profile-only", action="store_true",
                    help="few steps, no e2e / cpu baseline / rows (the command profiled under ncu)")
    args = ap.parse_args()
    if args.warmup < 3 and not args.profile_only:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import rd3_b200
    from rd3_b200 import _lib, parallel, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = synthetic.CONFIGS[WORKLOAD]
    H, W = cfg["hw"]
    BT = args.frames                                   # frames per step, whole job
    K, mv = cfg["max_points"], cfg["max_voxels"][0]
    npix = 6 * H * W
    C = F = 3

    # strong scaling (BASELINE config C5): the step's BT frames are sharded by sample, contiguous chunks
    f0, f1 = parallel.shard_range(BT, world, rank)
    B = f1 - f0
    sizes = parallel.shard_sizes(BT, world)
    bmax = max(sizes)
    host = synthetic.make_batch(list(range(f0, f1)), H, W, with_conf=not args.no_masks, scene=args.scene)
    depth_h = host["depth"].pin_memory()
    intr_h, c2l_h = host["intrinsics"].pin_memory(), host["cam2lidar"].pin_memory()
    depth, intr, c2l = depth_h.to(dev), intr_h.to(dev), c2l_h.to(dev)

    def make_mod(**kw):
        return rd3_b200.DepthToVoxels(cfg["voxel_size"], cfg["pcr"], K, cfg["max_voxels"], max_depth=synthetic.MAX_DEPTH,
                                      reuse_buffers=True, flat_outputs=world > 1, **kw).to(dev).train()

    mod = make_mod()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the step ----------------------------------------------------------------------
    # N == 1: one pass of the fused path over the BT frames.
    # N  > 1: the rank's shard in two sub-batches; what the sparse encoder consumes (voxel mean, coors, counts, voxel_num;
    #         fixed-size padded rows) is all-gathered on a communication stream while the next sub-batch -- of this step or
    #         of the next one -- computes.  Two sets of output / gather buffers alternate from step to step, so a step's
    #         gather only has to be finished when its buffers come up again two steps later; PIPE steps are replayed as
    #         one CUDA graph.
    PIPE = 4
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    # two sub-batches overlap a step's own gather with its own kernels; with few frames per GPU a call is latency-bound
    # (~14 dependent launches) and one sub-batch per step is faster -- its gather then overlaps the NEXT step's kernels
    nsub = 2 if (world > 1 and B >= 16) else 1
    subs = [(B * i // nsub, B * (i + 1) // nsub) for i in range(nsub)]
    sub_mods = [[make_mod() for _ in subs] for _p in range(2)] if world > 1 else None
    gathered = None
    if world > 1:
        # one flat int32 buffer per sub-batch and rank: [mean | coors | num | voxel_num] (DepthToVoxels(flat_outputs=True)),
        # gathered with ONE collective; the views below are what a consumer reads
        def flat_len(nb):
            return nb * mv * 7 + nb
        gathered_flat = [[torch.empty((world, flat_len(s1 - s0)), dtype=torch.int32, device=dev) for s0, s1 in subs]
                         for _p in range(2)]
        gathered = [dict(voxel_num=gf[:, (s1 - s0) * mv * 7:(s1 - s0) * mv * 7 + (s1 - s0)])
                    for gf, (s0, s1) in zip(gathered_flat[0], subs)]
        if len(set(sizes)) != 1:
            raise RuntimeError("--frames must be a multiple of the number of GPUs (equal shards are gathered without padding)")

    outs = [[None] * nsub, [None] * nsub]                  # the sub-batches' (reused) output buffers, per buffer set
    gdone = [None, None]                                   # per buffer set: its last gather has finished

    def compute_only():
        if world == 1:
            return mod(depth, intr, c2l)
        for si, ((s0, s1), m) in enumerate(zip(subs, sub_mods[0])):
            outs[0][si] = m(depth[s0:s1], intr[s0:s1], c2l[s0:s1])
        return outs[0][-1]

    def step_p(p):
        """One step on buffer set p: sub-batch kernels on the current stream, each followed by its all-gather on the
        communication stream.  Nothing waits for the gathers here."""
        cur = torch.cuda.current_stream(dev)
        if gdone[p] is not None:
            cur.wait_event(gdone[p])                       # the set's previous gather has read the buffers written below
        for si, ((s0, s1), m) in enumerate(zip(subs, sub_mods[p])):
            outs[p][si] = m(depth[s0:s1], intr[s0:s1], c2l[s0:s1])
            ev = torch.cuda.Event()
            ev.record(cur)
            with torch.cuda.stream(comm):
                comm.wait_event(ev)
                dist.all_gather_into_tensor(gathered_flat[p][si], outs[p][si]["flat"])
        gdone[p] = torch.cuda.Event()
        gdone[p].record(comm)
        return outs[p][-1]

    def join():
        torch.cuda.current_stream(dev).wait_stream(comm)

    step_count = [0]

    def step():
        if world == 1:
            return mod(depth, intr, c2l)
        r_ = step_p(step_count[0] & 1)
        step_count[0] += 1
        return r_

    for _ in range(args.warmup):
        r = step()
    if world > 1:
        join()
    barrier()
    graph = None
    if world > 1 and not args.no_graph:
        # 8 frames per GPU are ~40 kernel launches + 2 collectives for ~0.2 ms of device work: PIPE steps are replayed as
        # one CUDA graph (the library's internal stream lanes and NCCL are both capturable); eager if capture fails
        try:
            gs = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(gs):
                for k in range(PIPE):
                    step_p(k & 1)
                join()
            torch.cuda.synchronize()
            gdone[0] = gdone[1] = None                         # no event from outside the capture
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=gs):
                for k in range(PIPE):
                    step_p(k & 1)
                join()
            gdone[0] = gdone[1] = None
            g.replay()
            torch.cuda.synchronize()
            graph = g
        except Exception as e:                                 # noqa: BLE001
            sys.stderr.write("CUDA graph capture of the step failed (%s): eager launches\n" % (str(e).splitlines()[0],))
            graph = None
            gdone[0] = gdone[1] = None
            torch.cuda.synchronize()
        ok = torch.tensor([1 if graph is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            graph = None

    def run_steps(k):
        """exactly k steps, gathers included (the stream is joined with the communication stream at the end)"""
        if world == 1:
            for _ in range(k):
                mod(depth, intr, c2l)
            return
        if graph is not None:
            for _ in range(k // PIPE):
                graph.replay()
            k = k % PIPE
        for _ in range(k):
            step()
        join()

    run_steps(3 if world == 1 else PIPE + 1)
    barrier()
    vn = r["voxel_num"].tolist() if world == 1 else [v for gd in gathered for v in gd["voxel_num"].flatten().tolist()]
    M_total = int(sum(vn))                               # voxels of the whole job per step

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    # ---- device-resident timed region ------------------------------------------------
    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(0.3)
    ms_per_step = timed(lambda: run_steps(args.steps), 1) / args.steps
    clk = clocks.stop()
    value = BT * npix / (ms_per_step * 1e-3)
    multi = None
    if world > 1:
        n2 = max(3, min(args.steps, 30))
        compute_ms = timed(compute_only, n2)

        def gather_only():
            for si in range(nsub):
                dist.all_gather_into_tensor(gathered_flat[0][si], outs[0][si]["flat"])
        gather_ms = timed(gather_only, n2)
        gbytes = sum(t.numel() * t.element_size() for t in gathered_flat[0])
        # weak scaling for reference: BT frames on EVERY GPU, no collective (what round 1 reported)
        wh = synthetic.make_batch([rank * BT + i for i in range(BT)], H, W, with_conf=False, scene=args.scene)
        wd, wi, wc = wh["depth"].to(dev), wh["intrinsics"].to(dev), wh["cam2lidar"].to(dev)
        wmod = make_mod()
        for _ in range(3):
            wmod(wd, wi, wc)
        weak_ms = timed(lambda: wmod(wd, wi, wc), n2)
        del wd, wmod
        multi = {"frames_total": BT, "frames_per_gpu": B, "compute_only_ms": compute_ms, "gather_only_ms": gather_ms,
                 "compute_plus_gather_ms": ms_per_step, "gathered_bytes_per_rank_per_step": gbytes,
                 "cuda_graph": graph is not None,
                 "gather": "one all_gather_into_tensor per sub-batch of the flat [voxel_mean | coors | num_points | voxel_num] "
                           "buffer (padded rows, 28 B per voxel slot) on a communication stream, overlapped with the kernels of "
                           "the next sub-batch (of this step or the next: two buffer sets, %d steps per graph replay)" % PIPE,
                 "weak": {"frames_per_gpu": BT, "ms_per_step": weak_ms, "value": world * BT * npix / (weak_ms * 1e-3),
                          "frames_per_sec": world * BT / (weak_ms * 1e-3), "note": "independent replicas, no collective"}}

    # per-kernel durations: extra steps with CUDA events between the kernels (rd3_profile_*; the library then runs
    # its frame sub-batches on ONE stream so that the events bracket single kernels)
    prof_steps = max(1, min(args.steps, 5))
    _lib.profile_enable(True)
    for _ in range(prof_steps):
        compute_only()
    torch.cuda.synchronize()
    stage_ms, calls = _lib.profile_read()
    _lib.profile_enable(False)
    stage_ms = {k: v / prof_steps for k, v in stage_ms.items()}     # per step, this rank's shard

    # ---- roofline of the dominant kernel + of the whole fused path ---------------------
    hbm_peak, peak_src = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        pass
    M_local = M_total * B // BT
    # ALGORITHMIC bytes per launch (DESIGN.md "Kernels"): compulsory reads+writes only
    alg = {
        "insert": B * npix * 4,                                 # depth read (rounds of the still open frames)
        "lookup": B * npix * 4,                                 # depth read
        "flags": 0, "cull": 0, "memset": 0,                     # scratch-only stages
        "emit": M_local * (K * C * 4 + 12 + 4 + 4 * F),         # voxels + coors + num + mean written
        "meta": 0,
    }
    path_bytes = BT * npix * 4 + M_total * (K * 4 * C + 16) + M_total * 4 * F
    kern = {k: v for k, v in stage_ms.items() if k not in ("memset", "meta")}
    dom = max(kern, key=kern.get) if calls else "lookup"
    rounds = _lib.lib().rd3_hard_voxel_rounds(npix, max(1, B // (nsub if world > 1 else 1)))
    lanes = max(1, min(int(os.environ.get("RD3_STREAMS", "2")), 4, B))
    dom_launches = rounds if dom == "insert" else 1
    dom_ms = stage_ms[dom] / dom_launches
    dom_achieved = alg[dom] / dom_launches / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    path_achieved = path_bytes / (ms_per_step * 1e-3) / 1e9
    kname = {"insert": "hv_pass_kernel<DepthSource,0>", "lookup": "hv_pass_kernel<DepthSource,1>"}.get(dom, "hv_%s_kernel" % dom)
    roofline = {"bound": "hbm", "kernel": kname, "achieved": dom_achieved, "peak": hbm_peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": dom_achieved / hbm_peak, "traffic": None,
                "ms_per_launch": dom_ms, "launches_per_step": dom_launches,
                "algorithmic_bytes_per_launch": alg[dom] / dom_launches,
                "how": "CUDA events between the kernels of %d extra steps with the frame sub-batches on one stream" % prof_steps}
    if dom == "insert":
        roofline["note"] = ("the insert pass is %d launches over consecutive pixel ranges; a frame whose earlier rounds claimed "
                            "max_voxels voxels is closed and its CTAs of the later rounds return at once, so the average "
                            "launch is much shorter than the first one" % rounds)
    # second opinion: the other two heavy kernels on the same scale (algorithmic bytes of one launch / its duration)
    roofline["others"] = {k: {"ms_per_step": stage_ms[k], "achieved": alg[k] / (stage_ms[k] * 1e-3) / 1e9 if stage_ms[k] > 0 else 0.0,
                              "frac": alg[k] / (stage_ms[k] * 1e-3) / 1e9 / hbm_peak if stage_ms[k] > 0 else 0.0}
                          for k in ("insert", "lookup", "emit") if k != dom}
    traffic_file = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(traffic_file):
        try:
            tr_ = json.load(open(traffic_file))
            # an ncu number belongs to ONE build: it is only reported when the kernel sources are unchanged
            ent = (tr_.get("kernels") or {}).get(kname)
            if ent and tr_.get("source_hash") == _source_hash() and world == 1:
                roofline["traffic"] = ent.get("dram_bytes_per_launch")
                roofline["traffic_source"] = tr_.get("capture")
        except Exception:
            pass
    path_roofline = {"bound": "hbm", "achieved": path_achieved, "peak": hbm_peak * world, "unit": "GB/s",
                     "frac": path_achieved / (hbm_peak * world), "algorithmic_bytes_per_step": path_bytes,
                     "stage_ms_per_step_single_stream": stage_ms}

    # ---- the path WITH the confidence / sky masks and the device-side percentile threshold (north_star; SURVEY 8(d):
    #      42.1 MB / frame) -- same frames, conf + sky resident in HBM
    masks = None
    if not args.no_masks and not args.profile_only:
        conf, sky = host["conf"].to(dev), host["sky"].to(dev)
        mmod = make_mod()

        def mstep():
            thr = rd3_b200.conf_threshold(conf, sky, synthetic.CONF_PERCENTILE)
            return mmod(depth, intr, c2l, confs=conf, conf_thresh=thr, sky_masks=sky)

        for _ in range(3):
            rm = mstep()
        n3 = max(3, min(args.steps, 30))
        m_ms = timed(mstep, n3)
        Mm = torch.tensor([int(rm["voxel_num"].sum())], device=dev)
        if world > 1:
            dist.all_reduce(Mm)
        m_bytes = BT * npix * 9 + int(Mm.item()) * (K * 4 * C + 16 + 4 * F)
        masks = {"ms_per_step": m_ms, "frames_per_sec": BT / (m_ms * 1e-3), "value": BT * npix / (m_ms * 1e-3),
                 "algorithmic_bytes_per_step": m_bytes, "achieved_GBps": m_bytes / (m_ms * 1e-3) / 1e9,
                 "frac": m_bytes / (m_ms * 1e-3) / 1e9 / (hbm_peak * world), "voxels_per_frame_mean": int(Mm.item()) / BT,
                 "what": "conf >= percentile(conf[~sky], %g) (exact device-side selection, no host read) & ~sky & depth "
                         "masks fused into the same kernels; no collective" % synthetic.CONF_PERCENTILE}
        del conf, sky, mmod

    # ---- e2e: public API from pinned host buffers, copies inside the timed region ------
    e2e = None
    if not args.no_e2e and not args.profile_only:
        chunk = args.e2e_chunk if B % args.e2e_chunk == 0 else B
        nchunks = B // chunk
        nstreams = min(args.e2e_streams, nchunks)
        streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
        d_in = [torch.empty((chunk, 6, H, W), device=dev) for _ in range(nstreams)]

        def run_e2e(with_voxels):
            mods = [make_mod(with_voxels=with_voxels) for _ in range(nstreams)]
            if with_voxels:
                h_out = dict(voxels=torch.empty((B, mv, K, 3), pin_memory=True),
                             coors=torch.empty((B, mv, 3), dtype=torch.int32, pin_memory=True),
                             num_points=torch.empty((B, mv), dtype=torch.int32, pin_memory=True),
                             voxel_mean=torch.empty((B, mv, 3), pin_memory=True),
                             voxel_num=torch.empty((B,), dtype=torch.int32, pin_memory=True))
            else:       # what the sparse encoder consumes: packed (sum M, 3) features, (sum M, 4) [b,z,y,x] coors, counts
                h_out = dict(feats=torch.empty((nchunks, chunk * mv, 3), pin_memory=True),
                             coors=torch.empty((nchunks, chunk * mv, 4), dtype=torch.int32, pin_memory=True),
                             num=torch.empty((nchunks, chunk * mv), dtype=torch.int32, pin_memory=True),
                             offsets=torch.empty((nchunks, chunk + 1), dtype=torch.int32, pin_memory=True))
            h2d = depth_h.numel() * 4 + intr_h.numel() * 4 + c2l_h.numel() * 4
            d2h = sum(v.numel() * v.element_size() for v in h_out.values())

            def e2e_step():
                for ci in range(nchunks):
                    s = streams[ci % nstreams]
                    sl = slice(ci * chunk, (ci + 1) * chunk)
                    with torch.cuda.stream(s):
                        d_in[ci % nstreams].copy_(depth_h[sl], non_blocking=True)
                        k_d = intr_h[sl].to(dev, non_blocking=True)
                        m_d = c2l_h[sl].to(dev, non_blocking=True)
                        rr = mods[ci % nstreams](d_in[ci % nstreams], k_d, m_d)
                        if with_voxels:
                            for name, hbuf in h_out.items():
                                hbuf[sl].copy_(rr[name], non_blocking=True)
                        else:
                            of, oc, on, offs = rd3_b200.pack_sparse_inputs(rr, batch_offset=f0 + ci * chunk, with_num_points=True,
                                                                           sync=False, reuse_buffers=True)
                            # packed rows are a prefix of the worst-case buffers; the copy moves the worst case (the
                            # synthetic frames saturate max_voxels), the consumer reads offsets[-1] rows
                            h_out["feats"][ci].copy_(of, non_blocking=True)
                            h_out["coors"][ci].copy_(oc, non_blocking=True)
                            h_out["num"][ci].copy_(on, non_blocking=True)
                            h_out["offsets"][ci].copy_(offs, non_blocking=True)
                for s in streams:
                    s.synchronize()

            n_e2e = max(2, min(args.steps, 5))
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                e2e_step()
            barrier()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            res = {"value": BT * npix * n_e2e / float(dt.item()), "unit": UNIT,
                   "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": n_e2e,
                   "frames_per_sec": BT * n_e2e / float(dt.item())}
            if with_voxels:
                assert torch.equal(h_out["voxel_num"], compute_only()["voxel_num"].cpu()) or world > 1
            else:
                assert int(h_out["offsets"][0, -1]) > 0
            return res

        e2e = run_e2e(False)
        e2e["how"] = ("pinned host depth/calibration -> %d-frame chunks on %d streams: H2D, fused kernels without the padded "
                      "voxel tensor, pack_sparse_inputs, D2H of the packed features + [b,z,y,x] coors + counts + offsets "
                      "(what SparseEncoder consumes) through rd3_b200.DepthToVoxels; per-rank byte counts" % (chunk, nstreams))
        e2e["full_voxel_tensor"] = run_e2e(True)
        e2e["full_voxel_tensor"]["how"] = "same, D2H of voxels + coors + num + mean + count (padded tensors)"

    # ---- operator rows at the other BASELINE configs (C1, C3, C4; one GPU) ---------------
    rows = None
    if rank == 0 and world == 1 and not args.no_rows and not args.profile_only:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_rows
            rows = bench_rows.collect(iters=10, scene=args.scene)
        except Exception as e:                                 # noqa: BLE001
            rows = {"error": str(e).splitlines()[0] if str(e) else repr(e)}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) --------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.profile_only:
        ref = CpuReference(WORKLOAD, workers=args.cpu_workers)
        ref.step()
        tt, ff = 0.0, 0
        while tt < 10.0 and ff < 8 * ref.workers:
            w, f, _o = ref.step()
            tt += w
            ff += f
        ref.close()
        cpu_baseline = {"value": ff * npix / tt, "unit": UNIT, "cores": ref.workers, "kind": ref.kind,
                        "frames_per_sec": ff / tt, "sample": ref.describe() + ", %d frames in %.1f s" % (ff, tt)}

    if rank == 0:
        # launches of this repo's kernels per step and rank: per stream lane init + insert rounds + count + post (first
        # points, cull bits) + lookup + emit, plus the calibration kernel of every DepthToVoxels call
        calls_per_step = 1 if world == 1 else nsub
        per_call = 1 + min(lanes, max(1, B // calls_per_step)) * (rounds + 5)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "frames_per_sec": BT / (ms_per_step * 1e-3),
            "config": {"workload": WORKLOAD_DESC,
                       "scene": args.scene, "frames_per_step": BT, "frames_per_gpu_per_step": B, "pixels_per_frame": npix,
                       "voxels_per_frame_mean": M_total / BT,
                       "l2": "inputs larger than L2 (%.0f MB depth per step per GPU vs 126 MB L2)" % (B * npix * 4 / 1e6)
                             if B * npix * 4 > 126e6 else "per-GPU inputs (%.0f MB) fit L2: the outputs written per step (%.0f MB) "
                             "and the scratch do not" % (B * npix * 4 / 1e6, M_local * (K * 12 + 28) / 1e6),
                       "parallelism": ("BASELINE config C5: %d frames per step sharded by sample over %d GPU(s), %d per GPU; "
                                       "NCCL all-gather of the encoder inputs inside the timed region" % (BT, world, B))
                                      if world > 1 else "one GPU, %d frames per step, no collective" % BT},
            "clocks": clk, "gpu_launches": world * calls_per_step * per_call * args.steps,
            "roofline": roofline, "path_roofline": path_roofline,
            "e2e": e2e, "cpu_baseline": cpu_baseline,
        }
        if multi is not None:
            line["multi_gpu"] = multi
        if masks is not None:
            line["with_masks"] = masks
        if rows is not None:
            line["rows"] = rows
        _emit(line)
    if world > 1:
        _shutdown_process_group(dist)
    return 0


if __name__ == "__main__":
    sys.exit(main())
