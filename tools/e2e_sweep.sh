for mb in 2 4 8 16 32; do
  AWS_HUFFMAN_BATCH_SHARD_MB=$mb python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('shard_mb', $mb, 'e2e %.1f GB/s %.2f ms'%(j['e2e']['value'], j['e2e']['ms_per_step']))"
done
