#!/bin/bash
# compute-sanitizer memcheck over smoke() and the GPU suite (the 1 GiB tests excluded: minutes each under the sanitizer)
TAG=${1:-r02}; mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_memcheck.txt
{ echo "# compute-sanitizer --tool memcheck (B200), $(date -u +%F), library $(sha256sum stenos_b200/libstenos_b200.so | cut -c1-16)"; } > $OUT
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_memcheck_smoke.log 2>&1; echo "smoke(): rc $? | $(grep -E 'ERROR SUMMARY' gpurun_out/${TAG}_memcheck_smoke.log | tail -1)" | tee -a $OUT
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size and not reference_own and not holds_most" > gpurun_out/${TAG}_memcheck_tests.log 2>&1; echo "pytest -m gpu -k 'not full_size and not reference_own and not holds_most': rc $? | $(grep -E 'passed|failed' gpurun_out/${TAG}_memcheck_tests.log | tail -1) | $(grep -E 'ERROR SUMMARY' gpurun_out/${TAG}_memcheck_tests.log | tail -1)" | tee -a $OUT
grep -E "Invalid|out of bounds|misaligned" gpurun_out/${TAG}_memcheck_smoke.log gpurun_out/${TAG}_memcheck_tests.log | head -20 | tee -a $OUT
