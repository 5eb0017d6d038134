#!/bin/bash
# A/B helper: tools/ab.sh "ENV=1 ENV2=2" ... -> one timing line per environment string
for e in "$@"; do
  echo "== $e"
  env $e timeout 200 python tools/time_passes.py rtcamp6 1920 1080 9 2>&1 | tail -4
done
