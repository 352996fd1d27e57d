#!/bin/bash
# per-kernel durations of one step at SHAPE / BATCH (env), under ncu; table via /tmp/ktab.py-style parsing
REGEX=${REGEX:-'voxelize|tile_reduce|scan_buckets|place_|order_|bwd_plan|bwd_gather|bwd_pixel|fwd_dense|fwd_heavy'}
python profiles/fwd_variant_probe.py > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"$REGEX" -s ${SKIP:-46} -c ${COUNT:-11} --csv --log-file gpurun_out/launches_shape.csv python profiles/fwd_variant_probe.py > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/ncu.log
