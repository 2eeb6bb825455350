V=$PWD/aws-c-compression_b200/lib/variants
R2_WORKLOADS="hpack_batch" bash tools/r2_iter.sh pull2 --notest "X=1" "AWS_HUFFMAN_B200_LIB=$V/pull8.so" "AWS_HUFFMAN_B200_LIB=$V/pull6.so" "X=2"
