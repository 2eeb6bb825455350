#!/bin/bash
# A/B of library builds on one GPU box: tools/ab_variants.sh <tag> <variant> [<variant> ...]
# ("default" = lib/libaws-c-compression.so, anything else = lib/variants/<name>.so). Prints encode/decode ms
# of both workloads per variant and keeps the JSON lines in gpurun_out/<tag>_<variant>_<workload>.json.
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=""
  [ "$v" != default ] && lib=$PWD/aws-c-compression_b200/lib/variants/$v.so
  for w in hpack_batch stream; do
    AWS_HUFFMAN_B200_LIB=$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $w ${AB_ARGS} \
        > gpurun_out/${tag}_${v}_${w}.json 2> gpurun_out/${tag}_${v}_${w}.err
    python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/${tag}_${v}_${w}.json").read().strip().splitlines()[-1])
    print("%-10s %-12s enc %.3f ms  dec %.3f ms  value %.0f GB/s  e2e %.1f"%("$v","$w",j["encode_ms"],j["decode_ms"],j["value"],j["e2e"]["value"]))
except Exception as e:
    print("$v $w FAILED", e); print(open("gpurun_out/${tag}_${v}_${w}.err").read()[-1500:])
PY
  done
done
