#!/bin/bash
# One gpurun call: GPU tests, variant sweep, bench, ncu launch list and full captures.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
STEP=${1:-all}
if [ "$STEP" = all ] || [ "$STEP" = test ]; then
  : > gpurun_out/pytest_gpu.log
  for f in tests/test_stencils_gpu.py tests/test_halo_gpu.py tests/test_cpp_boundary.py tests/test_boundaries.py; do   # one process per file: a fault cannot cascade
    timeout 900 python -m pytest $f -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=300 >> gpurun_out/pytest_gpu.log 2>&1
    echo "pytest $f exit $?" >> gpurun_out/pytest_gpu.log
  done
  grep -E "passed|failed|exit" gpurun_out/pytest_gpu.log | tail -8
fi
if [ "$STEP" = sanitize ]; then
  timeout 900 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/sanitize.log 2>&1
  tail -15 gpurun_out/sanitize.log
fi
if [ "$STEP" = tuneva ]; then
  timeout 600 python tools/tune.py va > gpurun_out/tune_va.txt 2>&1
  head -12 gpurun_out/tune_va.txt
fi
if [ "$STEP" = all ] || [ "$STEP" = tune ]; then
  timeout 600 python tools/tune.py > gpurun_out/tune.txt 2>&1
  tail -3 gpurun_out/tune.txt
fi
if [ "$STEP" = all ] || [ "$STEP" = bench ]; then
  timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err
  cat gpurun_out/bench.json
  timeout 300 python bench.py --stencil hori_diff --steps 200 --warmup 20 --no-extras > gpurun_out/bench_hd.json 2>> gpurun_out/bench.err
  timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
fi
if [ "$STEP" = all ] || [ "$STEP" = ncu ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/ncu_launches.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:va_ -s 3 -c 2 -f -o gpurun_out/prof_va \
      python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/ncu_va.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_hd.csv \
      python bench.py --stencil hori_diff --steps 5 --warmup 3 --no-extras > gpurun_out/ncu_launches_hd.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:hd_ -s 3 -c 2 -f -o gpurun_out/prof_hd \
      python bench.py --stencil hori_diff --steps 5 --warmup 3 --no-extras > gpurun_out/ncu_hd.log 2>&1
  ls -la gpurun_out
fi
# generic paths of the C++ tag (specs without a hand-written kernel): timings beside stencil::gpu<>, block-geometry /
# unroll / prefetch sweep (make -C tests/cpp incumbent fused_timing), one ncu capture of each fused kernel
if [ "$STEP" = generic ]; then
  timeout 120 tests/_build/b200_generic > gpurun_out/b200_generic.txt 2>&1; tail -n 1 gpurun_out/b200_generic.txt
  timeout 120 tests/_build/gpu_incumbent > gpurun_out/incumbent.txt 2>&1; cat gpurun_out/incumbent.txt
  timeout 120 tests/_build/fused_timing > gpurun_out/fused_timing.txt 2>&1; cat gpurun_out/fused_timing.txt
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mss_kernel -s 20 -c 2 -f -o gpurun_out/prof_fused \
      tests/_build/fused_timing > gpurun_out/ncu_fused.log 2>&1
fi
