SD_HEADS_PAIR=0 timeout 120 python profiles/heads_time.py 2>&1 | tail -2
SD_HEADS_PAIR=1 timeout 120 python profiles/heads_time.py 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_e_fullsize.py tests/test_gpu_b_tc.py -x -q 2>&1 | tail -5
