#!/bin/bash
# N-GPU validation: NCCL gather equality test + the strong-scaling bench line (C5) at N GPUs
cd "$(dirname "$0")/.."
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
[ -n "$SKIPTEST" ] || timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k nccl > gpurun_out/pytest_nccl.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_nccl.log
timeout ${BT:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-50} --warmup 5 $EXTRA > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.json"))
print(d["n_gpus"], d["ms_per_step"], d["frames_per_sec"], d["scaling"]); print(json.dumps(d.get("multi_gpu"))[:1500]); print("e2e", json.dumps(d.get("e2e"))[:600]); print("masks", json.dumps(d.get("with_masks"))[:300])
PY
