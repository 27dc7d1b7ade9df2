#!/bin/bash
# Multi-GPU call (gpurun --gpus N): the cross-process IPC halo test, then bench.py under torchrun like the driver launches it
N=${1:-2}; TAG=${2:-m2a}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/smi.log 2>&1
if [ -z "$NO_TESTS" ]; then
timeout 600 python -m pytest tests/test_ipc_halo.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "ipc test exit $?"; tail -5 $OUT/pytest.log
fi
for CFG in ${CONFIGS:-default}; do
  EXTRA=""; [ "$CFG" != "default" ] && EXTRA="--width ${CFG%x*} --height ${CFG#*x}"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-20} --warmup ${WARMUP:-5} $EXTRA ${BENCH_ARGS} > $OUT/bench_n${N}_$CFG.json 2> $OUT/bench_n${N}_$CFG.err
  echo "bench N=$N $CFG exit $?"; tail -1 $OUT/bench_n${N}_$CFG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d.get(k) for k in ('value','frames_per_s_native','ms_per_frame','gpu_launches','parity','in_order')})
print('kernels', {k:(round(v['ms']*1e3,1), round(v['frac'],3), [round(x*1e3,1) for x in v['ms_per_rank']]) for k,v in (d.get('kernels') or {}).items()})
print('e2e', d.get('e2e') and d['e2e']['value'], 'workload', d['config']['workload'])
"; tail -3 $OUT/bench_n${N}_$CFG.err
done
if [ -n "$REF_ARM" ]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 --ref-budget 12 > $OUT/ref_n$N.json 2> $OUT/ref_n$N.err; echo "ref arm exit $?"; cut -c1-400 $OUT/ref_n$N.json
fi
