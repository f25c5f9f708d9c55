#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list of the bench command, one --set full capture of the render kernel.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [skip-tests]
tag=${1:-rX}
mkdir -p gpurun_out
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
  tail -3 gpurun_out/${tag}_pytest.log
fi
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 1500 gpurun_out/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_lean -s 2 -c 1 -f -o gpurun_out/${tag}_render python tools/prof_render.py 4 > gpurun_out/${tag}_ncu_render.log 2>&1
# the bench's own launch: 100 views in one launch (traffic per launch for roofline.traffic)
timeout 900 ncu --set full --clock-control none -k regex:render_lean -s 1 -c 1 -f -o gpurun_out/${tag}_render_batch python tools/prof_render.py 100 --batch > gpurun_out/${tag}_ncu_render_batch.log 2>&1
ls -la gpurun_out | tail -8
