#!/bin/bash
# larger BASELINE configurations: parity tests for them + one bench line each (kept under profiles/)
tag=${1:-rX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_io.py -m gpu -x -q > gpurun_out/${tag}_pytest_configs.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest_configs.log
tail -5 gpurun_out/${tag}_pytest_configs.log
for wl in sparse512 hires512_1080p hires512_4k; do
  timeout 600 python bench.py --workload $wl --steps 2 --warmup 3 > gpurun_out/${tag}_bench_${wl}.json 2> gpurun_out/${tag}_bench_${wl}.err
  tail -c 700 gpurun_out/${tag}_bench_${wl}.json; echo; tail -3 gpurun_out/${tag}_bench_${wl}.err
done
