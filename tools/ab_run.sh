#!/bin/bash
# tools/ab_run.sh "name[:bench flags]" ...   (run under gpurun) -> gpurun_out/ab.txt: env-steps/s of each variants/librr_<name>.so
out=gpurun_out/ab.txt; : > $out
for spec in "$@"; do
  v=${spec%%:*}; flags=""; [[ "$spec" == *:* ]] && flags=${spec#*:}
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --lib $PWD/variants/librr_$v.so $flags 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('$spec', round(d['value']), round(d['ms_per_step'],3))
    except Exception: print('$spec', l.strip()[:200])
" >> $out
done
cat $out
