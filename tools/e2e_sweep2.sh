for cfg in "2 5 8" "3 6 8" "4 8 8" "3 6 4" "4 8 4" "1 4 8"; do
  set -- $cfg
  AWS_HUFFMAN_BATCH_PIPE_DEPTH=$1 AWS_HUFFMAN_BATCH_PIPE_LANES=$2 AWS_HUFFMAN_BATCH_SHARD_MB=$3 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('depth $1 lanes $2 shard_mb $3', 'e2e %.1f GB/s %.2f ms'%(j['e2e']['value'], j['e2e']['ms_per_step']))"
done
