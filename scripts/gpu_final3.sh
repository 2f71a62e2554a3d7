#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest lm"; timeout 900 python -m pytest tests/test_gpu_lm.py -x -q -m gpu -s > gpurun_out/f3_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "split-KV|passed|failed|Error" gpurun_out/f3_pytest.log | tail -5
echo "== perf split3"; timeout 200 python profiles/perf_lm_step.py > gpurun_out/f3_perf_split3.log 2>&1; cat gpurun_out/f3_perf_split3.log
echo "== perf split1"; ACB_LM_ATT_SPLIT=1 timeout 200 python profiles/perf_lm_step.py > gpurun_out/f3_perf_split1.log 2>&1; cat gpurun_out/f3_perf_split1.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/f3_smoke.log 2>&1; echo "smoke rc=$?"
echo "== bench"; timeout 600 python bench.py > gpurun_out/f3_bench.json 2> gpurun_out/f3_bench.err; echo "bench rc=$?"; cut -c1-1400 gpurun_out/f3_bench.json
