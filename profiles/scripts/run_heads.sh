SD_HEADS_CHAIN=0 python profiles/heads_time.py 2>&1 | tail -1
SD_HEADS_CHAIN=1 python profiles/heads_time.py 2>&1 | tail -1
python -m pytest tests/test_gpu_e_fullsize.py tests/test_gpu_b_tc.py tests/test_gpu_g_pimg.py -x -q 2>&1 | tail -3
