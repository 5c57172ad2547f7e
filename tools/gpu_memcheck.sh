#!/bin/bash
# compute-sanitizer memcheck over smoke() and the GPU tests that exercise this round's new code paths (the row stores of the flow encoder
# that run past a row's end inside shared memory, the spill path, the bucket mode, the generic element sizes, the hybrid decoder)
TAG=${1:-r02}; mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_memcheck.txt
{ echo "# compute-sanitizer --tool memcheck (B200), $(date -u +%F)"; } > $OUT
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_memcheck_smoke.log 2>&1; echo "smoke(): rc $? | $(grep -E 'ERROR SUMMARY' gpurun_out/${TAG}_memcheck_smoke.log | tail -1)" | tee -a $OUT
STENOS_MEMCHECK=1 timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "golden_small or edge_cases or room_dependent or spill_path or mixed_superblocks or many_tiny or bucket or marker_252 or corrupt" > gpurun_out/${TAG}_memcheck_tests.log 2>&1; echo "pytest subset: rc $? | $(grep -E 'passed|failed' gpurun_out/${TAG}_memcheck_tests.log | tail -1) | $(grep -E 'ERROR SUMMARY' gpurun_out/${TAG}_memcheck_tests.log | tail -1)" | tee -a $OUT
grep -E "Invalid|out of bounds|misaligned" gpurun_out/${TAG}_memcheck_smoke.log gpurun_out/${TAG}_memcheck_tests.log | head -20 | tee -a $OUT
