#!/bin/bash
# One gpurun call: parity tests, bench, then (only if both exited 0) ncu launch list + full captures.
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
{
  echo "== host"; nproc; free -g | head -2; nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv
  echo "== opencl probe"; ls /etc/OpenCL/vendors 2>&1; find / -name 'libnvidia-opencl*' 2>/dev/null | head; find / -iname '*pocl*' 2>/dev/null | head -3
} > $OUT/probe.log 2>&1
python -m pytest tests -m gpu -x -q -s > $OUT/pytest.log 2>&1
PT=$?
echo "pytest exit $PT"; tail -15 $OUT/pytest.log
[ $PT -ne 0 ] && exit $PT
python bench.py --steps 3 --warmup 3 --all-modes > $OUT/bench.json 2> $OUT/bench.err
BE=$?
echo "bench exit $BE"; cat $OUT/bench.json | cut -c1-3000; tail -5 $OUT/bench.err
[ $BE -ne 0 ] && exit $BE
python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_reference.json 2>> $OUT/bench.err
cat $OUT/bench_reference.json | cut -c1-1200
SHORT="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
$SHORT > $OUT/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $SHORT > $OUT/ncu_launches.log 2>&1
$SHORT > $OUT/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fit_kernel|post_kernel' -s 40 -c 4 -f -o $OUT/prof $SHORT > $OUT/ncu_full.log 2>&1
ls -la $OUT
