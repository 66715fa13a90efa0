# full GPU test suite, headline bench, large-N chain probe
python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02c_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','schedules_ms','breakdown_ms','gpu_launches']}, d['e2e'], d['roofline']['frac'])
PY
for n in 2048 8192; do
SD_CHAIN=0 python profiles/imagine_time.py $n 16 10 2>&1 | tail -1
SD_CHAIN=1 python profiles/imagine_time.py $n 16 10 2>&1 | tail -1
done
