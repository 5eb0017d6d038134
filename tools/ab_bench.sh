#!/bin/bash
# A/B helper on the real bench loop: tools/ab_bench.sh "ENV=1 ENV2=2" ... -> value + ms per step per environment string
for e in "$@"; do
  echo "== $e"
  env $e timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step']/d['config']['passes_per_step'], d['detail']['kernel_time_share'])"
done
