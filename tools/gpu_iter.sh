#!/bin/bash
# One development iteration on the GPU box: parity tests, both benches, optional ncu captures.
#   tools/gpu_iter.sh <tag> [ncu-kernel-regex-batch] [ncu-kernel-regex-stream]
tag=${1:-it}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_batch.json 2> gpurun_out/${tag}_batch.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload stream > gpurun_out/${tag}_stream.json 2> gpurun_out/${tag}_stream.err
python - <<PY
import json
for w in ("batch","stream"):
    try:
        j=json.loads(open("gpurun_out/${tag}_%s.json"%w).read().strip().splitlines()[-1])
        print(w, "enc %.3f ms %.0f GB/s | dec %.3f ms %.0f GB/s | value %.0f | e2e %.1f (%.2f ms)"%(j["encode_ms"],j["encode_gbs_per_gpu"],j["decode_ms"],j["decode_gbs_per_gpu"],j["value"],j["e2e"]["value"],j["e2e"]["ms_per_step"]))
    except Exception as e:
        print(w, "FAILED", e); print(open("gpurun_out/${tag}_%s.err"%w).read()[-2000:])
PY
if [ -n "$2" ]; then
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s 2 -c 2 -o gpurun_out/${tag}_batch -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
fi
if [ -n "$3" ]; then
  ncu --set full --clock-control none --import-source on -k regex:"$3" -s 3 -c 3 -o gpurun_out/${tag}_stream -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload stream --stream-bytes 268435456 > /dev/null 2>&1
fi
ls -la gpurun_out | tail -5
