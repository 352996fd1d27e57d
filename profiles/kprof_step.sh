#!/bin/bash
# usage: tools_kprof.sh <tag>  — per-kernel durations + instruction counts of one bench step (under ncu)
TAG=$1
/usr/local/graft/bin/gpurun --timeout 900 -- "python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'fwd_dense|bwd_|order_|place_kernel|scan_buckets|voxelize|tile_reduce' -s 30 -c 10 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu.log 2>&1; echo rc=\$?" 2>&1 | tail -2
python - $TAG <<'PY'
import csv, collections, sys
with open(f'/root/repo/gpurun_out/launches_{sys.argv[1]}.csv') as f:
    lines=[l for l in f if not l.startswith('==')]
rows=list(csv.DictReader(lines))
agg=collections.OrderedDict()
for row in rows:
    k=(row['ID'],row['Kernel Name'][:28])
    agg.setdefault(k,{})[row['Metric Name']]=row['Metric Value']
for k,v in agg.items():
    g=lambda m: float(v.get(m,'0').replace(',',''))
    print(f"{k[1]:28s} {g('gpu__time_duration.sum')/1000:8.1f} us  rd={g('dram__bytes_read.sum')/1e6:7.1f}MB wr={g('dram__bytes_write.sum')/1e6:7.1f}MB warps={g('sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f}% inst={g('smsp__inst_executed.sum')/1e6:7.2f}M issue={g('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f}%")
PY
