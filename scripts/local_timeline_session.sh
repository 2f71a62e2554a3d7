#!/bin/bash
# Run HERE (build container): makes the instrumented build (-DACB_TIMELINE), ships it to a B200 for the layer-0 timeline of
# the decode step at KV 1 and 751, then restores the normal build.  Never leave the instrumented .so in the tree: the
# dormant stamps cost ~0.15 ms per step.
set -eu
cd "$(dirname "$0")/.."
ACB_BUILD_TIMELINE=1 python -c "import audiocraft_b200.build as b; b.build(force=True)"
/usr/local/graft/bin/gpurun --timeout 600 -- 'bash scripts/gpu_lm_timeline.sh' || true
python -c "import audiocraft_b200.build as b; b.build(force=True)"
tail -12 gpurun_out/v5_timeline_kv1.log || true
