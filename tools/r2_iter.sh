#!/bin/bash
# Round-2 iteration on the GPU box: parity tests, then the batch bench with each listed environment setting.
#   tools/r2_iter.sh <tag> [--notest] "ENV=.. ENV=.." ...
tag=$1; shift
mkdir -p gpurun_out
if [ "$1" == "--notest" ]; then shift; else
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${tag}_pytest.txt
fi
i=0
for envs in "$@"; do
  i=$((i+1))
  for w in ${R2_WORKLOADS:-hpack_batch}; do
    env $envs python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $w ${AB_ARGS} \
        > gpurun_out/${tag}_${i}_${w}.json 2> gpurun_out/${tag}_${i}_${w}.err
    python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/${tag}_${i}_${w}.json").read().strip().splitlines()[-1])
    print("%-40s %-12s enc %.3f ms  dec %.3f ms  value %.0f GB/s  e2e %.1f"%("$envs","$w",j["encode_ms"],j["decode_ms"],j["value"],j["e2e"]["value"]))
except Exception as e:
    print("$envs $w FAILED", e); print(open("gpurun_out/${tag}_${i}_${w}.err").read()[-1500:])
PY
  done
done
