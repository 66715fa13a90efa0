set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/r02b_pytest.log
SD_FUSE_BWD=1 python profiles/bwd_tail_time.py 2>&1 | tail -1
SD_FUSE_BWD=2 python profiles/bwd_tail_time.py 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02b_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','schedules_ms','breakdown_ms','gpu_launches']}, d['e2e'])
PY
