#!/bin/bash
# One GPU-box call of round 2: STAGES is a space-separated list of
#   tests    pytest -m gpu (TESTS="..." selects files / -k expressions)
#   opencl   the reference's unmodified bmfr.cl on the B200 (scripts/opencl_reference.py), 1080p fp32 + fp16 tmp_data, 720p
#   bench    bench.py at the driver's settings (+ BENCH_ARGS), then every tuning variant bmfr_b200/libbmfr_b200_*.so
#   launches ncu launch list of the bench command        full   ncu --set full of the three FUSED kernels
TAG=${1:-r2a}
STAGES=${STAGES:-"tests bench"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.log 2>&1
for S in $STAGES; do
case $S in
tests)
  timeout ${TEST_TIMEOUT:-900} python -m pytest ${TESTS:-tests} -m gpu -x -q --durations=8 > $OUT/pytest.log 2>&1
  echo "pytest exit $?"; tail -25 $OUT/pytest.log ;;
opencl)
  timeout 600 python scripts/opencl_reference.py --out $OUT/reference_opencl_1080p_fp32.json > $OUT/opencl_1080p_fp32.log 2>&1; echo "opencl fp32 exit $?"; tail -42 $OUT/opencl_1080p_fp32.log | cut -c1-600
  timeout 600 python scripts/opencl_reference.py --half 1 --compare 0 --out $OUT/reference_opencl_1080p_fp16.json > $OUT/opencl_1080p_fp16.log 2>&1; echo "opencl fp16 exit $?"; tail -40 $OUT/opencl_1080p_fp16.log | cut -c1-300
  timeout 600 python scripts/opencl_reference.py --width 1280 --height 720 --compare 0 --out $OUT/reference_opencl_720p_fp32.json > $OUT/opencl_720p_fp32.log 2>&1; echo "opencl 720p exit $?"
  timeout 600 python scripts/opencl_reference.py --width 256 --height 160 --frames 8 --port-frames 8 --out $OUT/reference_opencl_small_vs_port.json > $OUT/opencl_small.log 2>&1; echo "opencl small exit $?"; tail -3 $OUT/opencl_small.log | cut -c1-1500 ;;
bench)
  timeout 900 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS} > $OUT/bench.json 2> $OUT/bench.err
  echo "bench exit $?"; python scripts/show_bench.py $OUT/bench.json; tail -5 $OUT/bench.err
  for V in bmfr_b200/libbmfr_b200_*.so; do
    [ -f "$V" ] || continue
    N=$(basename $V .so)
    BMFR_B200_LIB=$PWD/$V timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --sustain-seconds 0 > $OUT/bench_$N.json 2>> $OUT/bench.err
    echo "== variant $N"; python scripts/show_bench.py $OUT/bench_$N.json
  done ;;
refarm)
  timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref arm exit $?"; cut -c1-900 $OUT/bench_ref.json ;;
launches)
  SHORT="timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --sustain-seconds 0"
  $SHORT > $OUT/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'reproject_tma_kernel|reproject_kernel|fit_gram_kernel|fit_qr_kernel|post_tma_kernel|post_kernel' -s 240 -c 90 --csv --log-file $OUT/launches.csv $SHORT > $OUT/ncu_launches.log 2>&1
  echo "launches exit $?" ;;
full)
  SHORT="timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --sustain-seconds 0"
  $SHORT > $OUT/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-reproject_kernel|fit_gram_kernel|post_tma_kernel}" -s ${NCU_SKIP:-243} -c ${NCU_COUNT:-3} -f -o $OUT/prof $SHORT > $OUT/ncu_full.log 2>&1
  echo "full exit $?" ;;
esac
done
ls $OUT
