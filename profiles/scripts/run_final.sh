# final evidence of the round: headline bench, reference arm, ncu launch list of the bench command (after its clean run)
python bench.py --steps 20 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02d_bench_ref.json 2> gpurun_out/r02d_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02d_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','schedules_ms','breakdown_ms','gpu_launches']}, d['e2e'], d['roofline']['frac'])
PY
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-imagine-bwd --no-gpu-reference --no-encoder > gpurun_out/r02d_b2.log 2>&1; echo "bench2 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02d_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-imagine-bwd --no-gpu-reference --no-encoder > gpurun_out/r02d_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/summarize_launches.py gpurun_out/r02d_bench_launches.csv > gpurun_out/r02d_bench_launches_summary.txt 2>&1; head -12 gpurun_out/r02d_bench_launches_summary.txt
