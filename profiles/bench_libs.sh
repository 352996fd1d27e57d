#!/bin/bash
# usage: profiles/bench_libs.sh [lib ...] — in-loop phase times (bench.py, nothing cached, cold inputs) per library build
cd "$(dirname "$0")/.."
one() { python bench.py --steps 100 --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d['phases_ms']; print('%s step %.1f us rank %.1f fwd %.1f plan %.1f bwd %.1f' % (sys.argv[1], d['ms_per_step']*1e3, p['rank_prepare']*1e3, p['forward']*1e3, p['bwd_plan']*1e3, p['backward']*1e3))" "$1"; }
one default
for l in "$@"; do FUSIONOCC_B200_LIB=$PWD/$l one $l; done
