# usage: tools/e2e_probe.sh "ENV1=a ENV2=b" "ENV1=c" ...   (each argument = one configuration; prints e2e of the batch workload)
for cfg in "$@"; do
  for rep in 1 2; do
  env $cfg python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', 'e2e %.1f GB/s %.2f ms'%(j['e2e']['value'], j['e2e']['ms_per_step']))"
  done
done
