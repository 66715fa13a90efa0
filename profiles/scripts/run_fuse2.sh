for m in 1 3 9 11 5 15; do echo "SD_FUSE_BWD=$m"; SD_FUSE_BWD=$m python profiles/bwd_tail_time.py 2>&1 | tail -1; done
