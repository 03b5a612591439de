#!/bin/bash
# round-2 GPU call AC: full GPU suite on the final tree (after the gn_apply rewrite and the scalar-trajectory test rule)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2ac_gputest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2ac_gputest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2ac_gputest.log | tail -8
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ac_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2ac_smoke.log
