mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash tools/r2_iter.sh rows3 --notest "X=1"
bash tools/r2_ncu.sh rows3 "decode_rows|compact_rows" hpack_batch 2>&1 | grep -E "rows|compact|ncu-rep"
