#!/bin/bash
# weak-scaling series (default workloads) and the strong-scaling BASELINE configs on N GPUs of one box
N=${1:-8}
OUT=gpurun_out/scale_n$N
mkdir -p $OUT
run() {  # name, extra args
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $N --steps 3 --warmup 2 $2 2> $OUT/$1.err | grep -E '^\{' > $OUT/$1.json
  python - "$OUT/$1.json" "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"{sys.argv[2]:14s} N={d['n_gpus']} {d['config']['workload'][:22]:22s} {d['ms_per_frame']*1000:8.1f} us/frame  native {d['frames_per_s_native']:8.1f} fps  1080p-eq {d['value']:9.1f} fps")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
WHAT=${2:-all}
run weak_p2p ""
run weak_nccl "--exchange nccl"
if [ "$WHAT" = "all" ]; then
  run strong4k_p2p "--width 3840 --height 2160"
  [ "$N" = "8" ] && run strong8k_p2p "--width 7680 --height 4320"
fi
grep -hiE "BmfrError|Error:" $OUT/*.err | head -5
