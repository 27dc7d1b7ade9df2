#!/bin/bash
# parity tests, bench (fused + staged per-kernel), launch list, ncu --set full of the four FUSED kernels
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.log 2>&1
if [ -z "$NCU_ONLY" ]; then
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
PT=$?
echo "pytest exit $PT"; tail -15 $OUT/pytest.log
[ $PT -ne 0 ] && [ -z "$KEEP_GOING" ] && exit $PT
timeout 600 python bench.py --steps 3 --warmup 3 ${BENCH_ARGS:---all-modes} > $OUT/bench.json 2> $OUT/bench.err
BE=$?
echo "bench exit $BE"; python scripts/show_bench.py $OUT/bench.json; tail -5 $OUT/bench.err
[ $BE -ne 0 ] && exit $BE
for V in bmfr_b200/libbmfr_b200_*.so; do
  [ -f "$V" ] || continue
  N=$(basename $V .so)
  BMFR_B200_LIB=$PWD/$V timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/bench_$N.json 2>> $OUT/bench.err
  echo "== variant $N"; python scripts/show_bench.py $OUT/bench_$N.json
done
fi
[ -n "$NO_NCU" ] && exit 0
# one ncu pass per call, each after the same command has exited 0 without ncu: NCU_MODE=launches (default) or full
SHORT="timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
if [ "${NCU_MODE:-launches}" = "launches" ]; then
  $SHORT > $OUT/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'reproject_kernel|fit_qr_kernel|post_kernel|noise_tile' -s 160 -c 80 --csv --log-file $OUT/launches.csv $SHORT > $OUT/ncu_launches.log 2>&1
else
  $SHORT > $OUT/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'reproject_kernel|fit_qr_kernel|post_kernel' -s 60 -c 3 -f -o $OUT/prof $SHORT > $OUT/ncu_full.log 2>&1
fi
ls $OUT
