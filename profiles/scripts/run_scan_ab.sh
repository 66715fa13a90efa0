# A/B of the persistent scan hand-off modes on one box: 0 = five grid barriers per step, 1 = two barriers + flagged hand-offs
# (a third mode with NO grid barrier -- flagged h_pre / ssq / deter hand-offs, every CTA polling every producer -- was built,
#  parity-green, and measured slower: 1.81 ms vs 1.62 ms; removed)
for ll in 1 2; do SD_SCAN_LL=$ll timeout 300 python -m pytest tests/test_gpu_f_pscan.py tests/test_gpu_a_fp32.py tests/test_gpu_j_fullsize.py tests/test_gpu_c_bwd.py -x -q -m gpu 2>&1 | tail -1; done
for i in 1 2; do for ll in 0 1 2; do echo "SD_SCAN_LL=$ll"; SD_SCAN_LL=$ll timeout 100 python profiles/observe_dist.py 40; done; done
for ll in 1 2; do SD_SCAN_LL=$ll SD_TRACE=1 SD_TRACE_SCAN=1 timeout 120 python profiles/observe_time.py 2>&1 | grep "SD_TRACE_SCAN" | tail -4; done
