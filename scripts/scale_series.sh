#!/bin/bash
# scaling evidence at the north-star sizes (BASELINE.json configs[3], [4]); usage: scripts/scale_series.sh <ngpus> <tag>
# writes one JSON line per run into gpurun_out/scale_<tag>_*.json
N=$1; TAG=$2; shift 2
run() {  # name, extra args...
  name=$1; shift
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --sustained-s 0 --no-f32 "$@" 2> gpurun_out/scale_${TAG}_${name}.err | grep '^{' > gpurun_out/scale_${TAG}_${name}.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $N --steps 10 --warmup 3 --sustained-s 0 "$@" \
      2> gpurun_out/scale_${TAG}_${name}.err | grep '^{' > gpurun_out/scale_${TAG}_${name}.json
  fi
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/scale_${TAG}_${name}.json"))
    print("${name}", "N=$N", "ms/step", round(d["ms_per_step"], 3), "pairs/s", f'{d["value"]:.3e}', "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"], 3), d.get("parity_check", {}).get("ok"))
except Exception as e:
    print("${name} FAILED", e)
PY
}
for job in "$@"; do
  case $job in
    weak) run weak_1e7 ;;
    strong) run strong_1e8 --scaling strong --n-total 1e8 --no-cpu ;;
    weak1e9) run weak_1.25e8 --n-per-gpu 1.25e8 --no-e2e --no-cpu ;;
    presorted) run presorted_1e7 --layout presorted --no-cpu ;;
    wide) run wide100_1e7 --box-xy 100 --no-cpu --no-e2e ;;
  esac
done
