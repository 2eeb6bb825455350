# usage: tools/ab_env.sh VAR "v1 v2 ..." [bench args]
var=$1; vals=$2; shift 2
for v in $vals; do
  env $var=$v python bench.py --steps 6 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$var=$v', 'enc %.3f dec %.3f ms'%(j['encode_ms'], j['decode_ms']))"
done
