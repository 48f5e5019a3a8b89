#!/bin/bash
# end-to-end rate of bench.py for several pool sizes: tools/e2e_workers.sh 8 12 16
for w in "$@"; do
  python bench.py --no-extras --no-cpu-baseline --no-ref-cuda --pool-workers $w 2>/dev/null | python -c '
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("workers", sys.argv[1], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1))' $w
done
