#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/f2_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/f2_pytest.log
echo "== perf ft32"; timeout 200 python profiles/perf_lm_step.py > gpurun_out/f2_perf_ft32.log 2>&1; cat gpurun_out/f2_perf_ft32.log
echo "== perf ft16"; ACB_LM_FT32=0 timeout 200 python profiles/perf_lm_step.py > gpurun_out/f2_perf_ft16.log 2>&1; cat gpurun_out/f2_perf_ft16.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/f2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/f2_smoke.log
echo "== bench"; timeout 600 python bench.py > gpurun_out/f2_bench.json 2> gpurun_out/f2_bench.err; echo "bench rc=$?"; cat gpurun_out/f2_bench.json
