#!/bin/bash
# gpurun --gpus 8: the timed 8K configuration with the default library and with the timing-only variants (no peer stores /
# no polls; results are wrong by construction) -> where the exchange's time goes.  Writes gpurun_out/$1/*.json
TAG=${1:-halo8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for V in ${VARIANTS:-default nopush nopoll}; do
  LIB=$PWD/bmfr_b200/libbmfr_b200.so; [ $V != default ] && LIB=$PWD/bmfr_b200/libbmfr_b200_$V.so
  [ -f $LIB ] || continue
  BMFR_B200_LIB=$LIB timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --parity-frames ${PARITY:-0} > $OUT/$V.json 2> $OUT/$V.err
  echo "== $V exit $?"; tail -1 $OUT/$V.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); t=d['timeline']
print('value', round(d['value']), 'in_order', round(d['in_order']['value']))
for k in ('period','reproject_busy','fit_busy','post_busy'): print(f'{k:16s}', t['connected'][k], 'alone', t['alone'][k][3])
"
done
