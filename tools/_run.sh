mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
V=$PWD/aws-c-compression_b200/lib/variants
R2_WORKLOADS="hpack_batch" bash tools/r2_iter.sh tma2 --notest "X=1" "AWS_HUFFMAN_B200_LIB=$V/notma.so"
