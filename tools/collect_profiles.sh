#!/bin/bash
# Collects the round's ncu evidence on the GPU box into gpurun_out/r1/ (summarised later with tools/ncu_summary.py).
mkdir -p gpurun_out/r1
K='hb::|encode_|decode_|stream_|scan_lens|tile_index|fill_packed|rebase|add_base'
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 400 --csv --log-file gpurun_out/r1/launches_hpack_batch.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 400 --csv --log-file gpurun_out/r1/launches_stream.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload stream > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"encode_slots_kernel|decode_batch" -s 2 -c 2 -o gpurun_out/r1/batch -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"encode_tiled|stream_fused_kernel" -s 2 -c 2 -o gpurun_out/r1/stream -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload stream --stream-bytes 268435456 > /dev/null 2>&1
ls -la gpurun_out/r1
# the bench lines of the same build (default arguments)
python bench.py > gpurun_out/r1/bench_hpack_batch.json 2> gpurun_out/r1/bench_hpack_batch.err
python bench.py --workload stream > gpurun_out/r1/bench_stream.json 2> gpurun_out/r1/bench_stream.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1/bench_reference.json 2> gpurun_out/r1/bench_reference.err
tail -c 600 gpurun_out/r1/bench_hpack_batch.json
# the widened rows: histogram and literal framing kernels
ncu --set full --clock-control none --import-source on -k regex:"histogram_kernel" -s 3 -c 1 -o gpurun_out/r1/histogram -f \
    python tools/histogram_probe.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"encode_|decode_|scan_lens|tile_index|fill_packed|hpack_" -c 200 --csv --log-file gpurun_out/r1/launches_literals.csv \
    python tools/literals_probe.py > /dev/null 2>&1
python tools/histogram_probe.py > gpurun_out/r1/histogram.json 2>/dev/null
python tools/literals_probe.py > gpurun_out/r1/literals.json 2>/dev/null
# BASELINE config 5's share of one GPU: 8M strings (64M over 8 GPUs)
python bench.py --strings 8000000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1/bench_8M_strings.json 2> gpurun_out/r1/bench_8M_strings.err
tail -c 400 gpurun_out/r1/bench_8M_strings.json; tail -3 gpurun_out/r1/bench_8M_strings.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
