#!/bin/bash
# One gpurun call: GPU parity tests (one process per file so a trapped kernel cannot poison the rest), bench line,
# ncu launch list + full captures.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 120 python scripts/debug_conv_tc.py > gpurun_out/debug_conv_tc.log 2>&1; echo "debug_conv_tc rc=$?"; tail -20 gpurun_out/debug_conv_tc.log
: > gpurun_out/pytest_gpu.log
for f in tests/test_ops_gpu.py tests/test_rick_gpu.py tests/test_conv_tc_gpu.py tests/test_model_gpu.py tests/test_adapt_gpu.py; do
  echo "=== $f" >> gpurun_out/pytest_gpu.log
  timeout 900 python -m pytest $f -m gpu -q -rA --timeout=300 2>&1 | tail -120 >> gpurun_out/pytest_gpu.log
done
grep -E "^(=== |FAILED|ERROR|[0-9]+ (passed|failed))|passed|failed" gpurun_out/pytest_gpu.log | tail -40
if [ "$1" != "tests-only" ]; then
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'upfirdn2d_tiled|bias_act' -c 14 \
    -o gpurun_out/prof_ops python scripts/prof_ops.py > gpurun_out/ncu_ops.log 2>&1
fi
ls -la gpurun_out
