for w in 0 1; do echo "SD_L2_WINDOW=$w"; SD_L2_WINDOW=$w python profiles/overlap_probe.py 2>&1 | tail -7; done
python -m pytest tests/test_gpu_c_bwd.py tests/test_gpu_f_pscan.py -x -q 2>&1 | tail -2
