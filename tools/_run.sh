mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_batch.py -m gpu -x -q -k "golden or rfc7541 or committed" > $O/sanitizer_racecheck.txt 2>&1
echo "racecheck exit $?" >> $O/sanitizer_racecheck.txt; tail -3 $O/sanitizer_racecheck.txt
for mb in 4 6 8 12 16; do
  AWS_HUFFMAN_BATCH_SHARD_MB=$mb python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --workload hpack_batch 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('shard_mb $mb e2e %.1f GB/s (%.2f ms)'%(j['e2e']['value'], j['e2e']['ms_per_step']))"
done
for cfg in "AWS_HUFFMAN_BATCH_PIPE_DEPTH=3" "AWS_HUFFMAN_BATCH_PIPE_DEPTH=3 AWS_HUFFMAN_BATCH_PIPE_LANES=8" "AWS_HUFFMAN_BATCH_PIPE_DEPTH=1 AWS_HUFFMAN_BATCH_PIPE_LANES=4"; do
  env $cfg python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --workload hpack_batch 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg e2e %.1f GB/s (%.2f ms)'%(j['e2e']['value'], j['e2e']['ms_per_step']))"
done
V=$PWD/aws-c-compression_b200/lib/variants
R2_WORKLOADS="stream" bash tools/r2_iter.sh pre --notest "X=1" "AWS_HUFFMAN_B200_LIB=$V/pre320.so" "AWS_HUFFMAN_B200_LIB=$V/pre448.so" "AWS_HUFFMAN_B200_LIB=$V/steps8.so" "AWS_HUFFMAN_B200_LIB=$V/steps4.so"
R2_WORKLOADS="hpack_batch" bash tools/r2_iter.sh stp --notest "AWS_HUFFMAN_B200_LIB=$V/steps8.so" "AWS_HUFFMAN_B200_LIB=$V/steps4.so"
