#!/bin/bash
# usage: scripts/sweep_opts.sh "opt=val,opt=val" ...   -> one quick device-resident bench line per option set
for o in "$@"; do
  MDB_OPTS="$o" python bench.py --steps 20 --warmup 4 --no-configs --no-next-rows --no-cpu-baseline --no-e2e --no-extra --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$o', 'fps %.0f  ms/step %.2f  chain alone %.3f  in-step frac %.3f  whole-step frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['chain_ms_per_batch'], d['roofline']['frac_inside_timed_region'], d['roofline']['frac_whole_step']))"
done
