#!/bin/bash
# compute-sanitizer passes over a few GPU parity tests (run under gpurun); logs in gpurun_out/
set -u
mkdir -p gpurun_out
T1='tests/test_gpu_parity.py::test_polish_matches_oracle_and_reference[0-m1]'
T2='tests/test_gpu_parity.py::test_polish_matches_oracle_and_reference[160-m2_i2_d3]'
T3='tests/test_gpu_binned.py::test_binned_scan_polish_matches_oracle[m1-chunks]'
T4='tests/test_gpu_binned.py::test_binned_scan_polish_matches_oracle[cbf_m1-overflow]'
T5='tests/test_gpu_binned.py::test_binned_scan_polish_matches_oracle[m1-regions]'
compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log python -m pytest "$T1" "$T2" "$T3" "$T4" -x -q
compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck.log python -m pytest "$T1" "$T3" -x -q
compute-sanitizer --tool synccheck --error-exitcode 9 --log-file gpurun_out/synccheck.log python -m pytest "$T5" -x -q
tail -2 gpurun_out/memcheck.log gpurun_out/racecheck.log gpurun_out/synccheck.log
