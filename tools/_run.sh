mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --workload hpack_batch 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default e2e %.1f GB/s (%.2f ms)'%(j['e2e']['value'], j['e2e']['ms_per_step']))"
