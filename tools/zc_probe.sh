for m in 1 2; do
  for w in hpack_batch stream; do
    AWS_HUFFMAN_BATCH_ZC_MODE=$m python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload $w 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('zc_mode $m $w e2e %.1f GB/s %.2f ms'%(j['e2e']['value'], j['e2e']['ms_per_step']))"
  done
done
