SD_PACK_TILED=0 python profiles/refresh_time.py 2>&1 | tail -1
SD_PACK_TILED=1 python profiles/refresh_time.py 2>&1 | tail -1
python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02d_pytest.log
