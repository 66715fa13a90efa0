run() {
  env "$@" python bench.py --steps 20 --warmup 3 --no-gpu-reference --no-encoder --no-imagine-bwd --no-cpu-baseline > gpurun_out/knob.json 2> gpurun_out/knob.err
  python - "$*" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/knob.json').read().strip().splitlines()[-1])
print(sys.argv[1], "| pass", round(d['ms_per_step'],3), "| sched", {k:round(v,3) for k,v in d['schedules_ms'].items()}, "| rollout lw", round(d['breakdown_ms']['imagine_fwd_layerwise'],3))
PY
}
run SD_X=0
run SD_TC_WIDE_BLOCK=1
run SD_TC_WIDE=1
run SD_TC_WIDE=1 SD_TC_WIDE_BLOCK=1
run SD_TC_SPLIT=0
run SD_PIMG_BG_TEAMS=2
run SD_PIMG_BG_TEAMS=3
