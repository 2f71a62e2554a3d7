#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/f5_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/f5_pytest.log
echo "== perf"; timeout 200 python profiles/perf_lm_step.py > gpurun_out/f5_perf.log 2>&1; cat gpurun_out/f5_perf.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/f5_smoke.log 2>&1; echo "smoke rc=$?"
echo "== bench"; timeout 600 python bench.py > gpurun_out/f5_bench.json 2> gpurun_out/f5_bench.err; echo "bench rc=$?"; cut -c1-1400 gpurun_out/f5_bench.json
