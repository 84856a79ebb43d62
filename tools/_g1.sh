mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_cli_gpu.py -m gpu -x -q > gpurun_out/r02aj_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02aj_pytest.txt
tail -3 gpurun_out/r02aj_pytest.txt
NTB_TUNE_E2E=1 python tools/scan_tune.py '' > gpurun_out/r02aj_c2.log 2> gpurun_out/r02aj_c2.err
cat gpurun_out/r02aj_c2.log
