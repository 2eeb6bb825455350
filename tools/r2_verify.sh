#!/bin/bash
# Final check on the GPU box: smoke(), the GPU suite, the default bench line (everything under a timeout).
mkdir -p gpurun_out/r2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r2/bench_verify.json 2> gpurun_out/r2/bench_verify.err; python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2/bench_verify.json").read().strip().splitlines()[-1])
print("value %.1f enc %.4f dec %.4f frac %.4f e2e %.1f | stream %.1f | sharded %.1f | parity %s %s"%(j["value"],j["encode_ms"],j["decode_ms"],j["roofline"]["frac"],j["e2e"]["value"],j["stream"]["value"],j["sharded_batch"]["value"],j["parity_checked"]["hpack_batch"]["encoded_bytes_equal"],j["parity_checked"]["stream"]["encoded_bytes_equal"]))
PY
