#!/bin/bash
# One gpurun call.  Every step has its own short timeout and writes its log unbuffered, so a hung kernel costs
# minutes, not the whole call.   usage: gpu_round.sh [tests] [bench] [ncu] [ncuops] [smoke]
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "cores: $(nproc)" >> gpurun_out/gpu.txt
for STEP in "$@"; do
case $STEP in
tests)
  for f in ${TEST_FILES:-tests/test_ops_gpu.py tests/test_rick_gpu.py tests/test_conv_tc_gpu.py tests/test_conv_train_gpu.py tests/test_model_gpu.py tests/test_adapt_gpu.py tests/test_ref_cuda_gpu.py}; do
    log=gpurun_out/pytest_$(basename $f .py).log
    timeout ${TEST_TIMEOUT:-420} python -u -m pytest $f -m gpu -q -rA --timeout=300 -p no:cacheprovider ${PYTEST_ARGS} > $log 2>&1
    echo "=== $f rc=$? : $(tail -1 $log)"
    grep -E "^(FAILED|ERROR)" $log | head -12
  done ;;
smoke)
  timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ;;
bench)
  timeout 600 python -u bench.py --steps ${BENCH_STEPS:-20} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench${BENCH_TAG}.json 2> gpurun_out/bench${BENCH_TAG}.err
  echo "bench rc=$?"; tail -c 6000 gpurun_out/bench${BENCH_TAG}.json; tail -5 gpurun_out/bench${BENCH_TAG}.err ;;
ncu)
  timeout 560 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv \
      --log-file gpurun_out/launches${NCU_TAG}.csv \
      python -u bench.py --steps 1 --warmup 3 --no-cpu-baseline --quick --cuda-profiler ${NCU_BENCH_ARGS:---mode eager} \
      > gpurun_out/ncu_bench${NCU_TAG}.log 2>&1
  echo "ncu launches rc=$?" ;;
ncutraffic)
  # DRAM bytes of every conv_tc / conv_wgrad launch of ONE timed iteration (one metric pass): roofline.traffic
  timeout 400 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --clock-control none -k regex:'conv_tc_kernel|conv_wgrad_kernel' -c 400 --csv --log-file gpurun_out/conv_traffic${NCU_TAG}.csv \
      python -u bench.py --steps 1 --warmup 3 --no-cpu-baseline --quick --cuda-profiler --mode eager \
      > gpurun_out/ncu_traffic${NCU_TAG}.log 2>&1
  echo "ncu traffic rc=$?" ;;
ncuops)
  timeout 300 ncu --set full --clock-control none --import-source on \
      -k regex:${NCU_KERNELS:-'upfirdn2d|bias_act_vec|bias_act_bwd|blur_nhwc|conv_tc|conv_wgrad|adam_mask_ema'} -s ${NCU_SKIP:-9} -c ${NCU_COUNT:-9} \
      -o gpurun_out/prof_ops${NCU_TAG} python -u ${NCU_SCRIPT:-scripts/prof_ops.py} > gpurun_out/ncu_ops${NCU_TAG}.log 2>&1
  echo "ncu ops rc=$?"
  # the raw metric page as CSV (small); the report itself only travels back when it fits the 64 MiB pull limit
  ncu -i gpurun_out/prof_ops${NCU_TAG}.ncu-rep --page raw --csv > gpurun_out/prof_ops${NCU_TAG}_raw.csv 2>/dev/null
  if [ -n "${NCU_DROP_REP}" ] || [ $(stat -c %s gpurun_out/prof_ops${NCU_TAG}.ncu-rep 2>/dev/null || echo 0) -gt 40000000 ]; then
    rm -f gpurun_out/prof_ops${NCU_TAG}.ncu-rep; fi ;;
esac
done
ls -la gpurun_out
