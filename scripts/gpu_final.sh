#!/bin/bash
# Round-end style validation: full GPU test suite, smoke, dependency-latency probe, bench (our arm).
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/final_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/final_smoke.log
echo "== chain latency probe"; timeout 300 python profiles/perf_chain_latency.py > gpurun_out/chain_latency.log 2>&1; cat gpurun_out/chain_latency.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; cat gpurun_out/final_bench.json
