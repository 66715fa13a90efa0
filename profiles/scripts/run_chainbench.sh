for ch in 0 1; do
SD_CHAIN=$ch python bench.py --steps 20 --warmup 3 --no-gpu-reference --no-encoder --no-imagine-bwd --no-cpu-baseline > gpurun_out/r02c_bench_chain$ch.json 2> gpurun_out/r02c_bench_chain$ch.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02c_bench_chain$ch.json').read().strip().splitlines()[-1])
print("SD_CHAIN=$ch", {k:d[k] for k in ['value','ms_per_step','schedules_ms']}, d['breakdown_ms']['imagine_fwd_layerwise'], d['e2e'])
PY
done
