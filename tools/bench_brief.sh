#!/bin/bash
# brief A/B line: steps/s + per-class device ms (no CPU baseline)
timeout 300 python bench.py --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('steps/s %.1f  e2e %.1f  ms/step %.4f  breakdown %s  clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], {k: round(v, 4) for k, v in d['breakdown_ms'].items()}, d['clocks']))"
