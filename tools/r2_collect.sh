#!/bin/bash
# Round-2 evidence on the GPU box -> gpurun_out/r2/ (summarised into profiles/ with tools/ncu_summary.py).
#   tools/r2_collect.sh [notest] [nobench] [noncu]
O=gpurun_out/r2; mkdir -p $O
K='str_|encode_|decode_|stream_|scan_lens|tile_index|fill_packed|rebase|add_base'
if [[ " $* " != *" notest "* ]]; then
  timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/pytest_gpu.txt
fi
if [[ " $* " != *" nobench "* ]]; then
  timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 1500 $O/bench_default.json; tail -5 $O/bench_default.err
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
fi
if [[ " $* " != *" noncu "* ]]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 400 --csv --log-file $O/launches_hpack_batch.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --workload hpack_batch > /dev/null 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 400 --csv --log-file $O/launches_stream.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --workload stream > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"str_pack|str_bits|str_scan|str_prep|decode_batch" -s 5 -c 5 -o $O/batch -f \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --workload hpack_batch > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"encode_tiled|stream_fused_kernel" -s 2 -c 2 -o $O/stream -f \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --workload stream > /dev/null 2>&1
  ls -la $O
fi
if [[ " $* " == *" sanitizer "* ]]; then
  # compute-sanitizer over a subset of the GPU parity suite (memcheck: every access; racecheck: shared-memory hazards)
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_batch.py -m gpu -x -q \
      -k "golden or rfc7541 or committed or random_batches or single_stream or empty" > $O/sanitizer_memcheck.txt 2>&1
  echo "memcheck exit $?" >> $O/sanitizer_memcheck.txt; tail -4 $O/sanitizer_memcheck.txt
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_batch.py -m gpu -x -q \
      -k "golden or rfc7541 or committed" > $O/sanitizer_racecheck.txt 2>&1
  echo "racecheck exit $?" >> $O/sanitizer_racecheck.txt; tail -4 $O/sanitizer_racecheck.txt
fi
