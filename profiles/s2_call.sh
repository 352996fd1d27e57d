set -x
mkdir -p gpurun_out
python profiles/step_ab.py --env FO_BWD_HALF=0 --env FO_BWD_HALF=1 > gpurun_out/s2_half_base8.json 2> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --batch 1 --rounds 1 --env FO_BWD_HALF=0 --env FO_BWD_HALF=1 > gpurun_out/s2_half_base1.json 2>> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --shape native --iters 100 --rounds 1 --env FO_BWD_HALF=0 --env FO_BWD_HALF=1 > gpurun_out/s2_half_native8.json 2>> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --shape stress --batch 2 --iters 100 --rounds 1 --env FO_BWD_HALF=0 --env FO_BWD_HALF=1 > gpurun_out/s2_half_stress2.json 2>> gpurun_out/s2_step_ab.err
cat gpurun_out/s2_half_*.json; tail -3 gpurun_out/s2_step_ab.err
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/s2_pytest2.log
for h in 0 1; do
FO_BWD_HALF=$h ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:"bwd_gather" --clock-control none -c 3 --csv --log-file gpurun_out/s2_half_ncu$h.csv python profiles/step_ab.py --rounds 1 --iters 1 > /dev/null 2>&1
grep -E "dram__bytes_read|gpu__time" gpurun_out/s2_half_ncu$h.csv | tail -2
done
