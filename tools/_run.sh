mkdir -p gpurun_out
V=$PWD/aws-c-compression_b200/lib/variants
for v in default pk3a pk3b pk2c default pk3a pk3b; do
  lib=""; [ "$v" != default ] && lib=$V/$v.so
  AWS_HUFFMAN_B200_LIB=$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload hpack_batch 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v enc %.4f ms dec %.4f ms parity %s'%(j['encode_ms'], j['decode_ms'], j['parity_checked']['hpack_batch']['encoded_bytes_equal']))"
done
