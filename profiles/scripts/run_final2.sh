# final evidence of the round (after the flagged hand-offs of the posterior scan): headline bench, reference arm, ncu launch list of
# the bench command (after its clean run), ncu --set full of the persistent scan kernel
python bench.py --steps 20 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02e_bench_ref.json 2> gpurun_out/r02e_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','schedules_ms','breakdown_ms','gpu_launches']}, d['e2e'], d['roofline']['frac'])
PY
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-imagine-bwd --no-gpu-reference --no-encoder > gpurun_out/r02e_b2.log 2>&1; echo "bench2 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02e_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-imagine-bwd --no-gpu-reference --no-encoder > gpurun_out/r02e_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/summarize_launches.py gpurun_out/r02e_bench_launches.csv > gpurun_out/r02e_bench_launches_summary.txt 2>&1; head -8 gpurun_out/r02e_bench_launches_summary.txt
python profiles/observe_scan_only.py 16 > gpurun_out/r02e_scan_only.log 2>&1; echo "scan_only rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:observe_scan -c 1 -o gpurun_out/r02e_scan python profiles/observe_scan_only.py 16 > gpurun_out/r02e_ncu_scan.log 2>&1; echo "ncu scan rc=$?"
python profiles/summarize_ncu.py gpurun_out/r02e_scan.ncu-rep > gpurun_out/r02e_ncu_full_observe_scan.txt 2>&1; cat gpurun_out/r02e_ncu_full_observe_scan.txt | head -24
