# persistent posterior scan after a change: parity tests that touch it, then stand-alone timing and the in-kernel phase stamps
timeout 600 python -m pytest tests/test_gpu_f_pscan.py tests/test_gpu_a_fp32.py tests/test_gpu_e_fullsize.py tests/test_gpu_j_fullsize.py tests/test_gpu_c_bwd.py -x -q -m gpu > gpurun_out/scan_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/scan_pytest.log
timeout 120 python profiles/observe_time.py 2>&1 | tail -3
SD_TRACE=1 SD_TRACE_SCAN=1 timeout 120 python profiles/observe_time.py 2>&1 | grep SD_TRACE_SCAN | tail -3
