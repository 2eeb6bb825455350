# A/B on the same box: AWS_HUFFMAN_BATCH_EXPERIMENT values given as arguments
for v in "$@"; do
  for rep in 1 2; do
  AWS_HUFFMAN_BATCH_EXPERIMENT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('exp', '$v', 'enc %.3f dec %.3f ms'%(j['encode_ms'], j['decode_ms']))"
  done
done
