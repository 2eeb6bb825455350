mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
V=$PWD/aws-c-compression_b200/lib/variants
R2_WORKLOADS="hpack_batch" bash tools/r2_iter.sh ab3 --notest "X=1" "AWS_HUFFMAN_BATCH_NO_SLOTS_DECODE=1"
bash tools/r2_ncu.sh slots2 "decode_slots|str_bits" hpack_batch 2>&1 | grep -E "slots|str_|ncu-rep"
