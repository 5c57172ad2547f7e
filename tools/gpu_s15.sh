#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 oracle/_ref/ref_comp_decomp_cut_b200 ) > gpurun_out/r2s15_ref_cd.log 2>&1; echo "ref comp_decomp rc $?"; tail -4 gpurun_out/r2s15_ref_cd.log
( time timeout 420 oracle/_ref/ref_test_cvector_main_b200 ) > gpurun_out/r2s15_ref_cv.log 2>&1; echo "ref cvector rc $?"; tail -6 gpurun_out/r2s15_ref_cv.log
( time timeout 300 oracle/_ref/cvector_dropin_b200 /tmp/cv.bin ) > gpurun_out/r2s15_dropin.log 2>&1; tail -5 gpurun_out/r2s15_dropin.log
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/r2s15_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2s15_pytest.log
tail -3 gpurun_out/r2s15_pytest.log
