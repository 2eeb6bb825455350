#!/bin/bash
# tools/r2_ncu.sh <tag> <kernel-regex> [workload] : launch list + one full capture of the named kernels
tag=$1; rx=$2; w=${3:-hpack_batch}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload $w ${AB_ARGS} > gpurun_out/${tag}_launch_bench.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/${tag}_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ki].split("(")[0]; v=float(r[vi].replace(",","")); u=r[ui]
    v = v/1000 if u in ("ns","nsecond") else v
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,t) in agg.items(): print("%-60s n=%3d total %9.1f us  avg %8.1f us"%(k[:60],c,t,t/c))
PY
if [ -n "$rx" ]; then
ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 2 -c 3 -o gpurun_out/${tag}_full -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload $w ${AB_ARGS} > /dev/null 2>&1
ls -la gpurun_out/${tag}_full.ncu-rep
fi
