# does the flagged-hand-off scan behave differently in a multi-process (torchrun + NCCL) run?  observe_fwd per mode, 2 ranks vs 1
for ll in 0 1 2; do
SD_SCAN_LL=$ll python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$ll bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-encoder --no-imagine-bwd > gpurun_out/llm_$ll.json 2> gpurun_out/llm_$ll.err
python - <<PY
import json
d=json.loads(open("gpurun_out/llm_$ll.json").read().strip().splitlines()[-1])
print("2 ranks SD_SCAN_LL=$ll", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["breakdown_ms"].items() if v})
PY
done
SD_SCAN_LL=2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-encoder --no-imagine-bwd > gpurun_out/llm_s.json 2> gpurun_out/llm_s.err
python - <<PY
import json
d=json.loads(open("gpurun_out/llm_s.json").read().strip().splitlines()[-1])
print("1 rank  SD_SCAN_LL=2", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["breakdown_ms"].items() if v})
PY
