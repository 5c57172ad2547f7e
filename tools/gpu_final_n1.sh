#!/bin/bash
# final N=1 evidence of a session: bench line (both arms), ncu summaries (tools/gpu_evidence.sh); "tests" as 2nd argument runs the GPU suite first
TAG=${1:-r02f}; mkdir -p gpurun_out
if [ "$2" = tests ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
fi
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc $?"; tail -3 gpurun_out/${TAG}_bench_n1.err; cut -c1-1500 gpurun_out/${TAG}_bench_n1.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref_n1.json 2> gpurun_out/${TAG}_bench_ref_n1.err; echo "ref rc $?"; cut -c1-600 gpurun_out/${TAG}_bench_ref_n1.json
tools/gpu_evidence.sh ${TAG}
