set -x
mkdir -p gpurun_out
python profiles/step_ab.py --env FO_BWD_RIDE=1 --env FO_BWD_RIDE=0 --env FO_BWD_RIDE=2 > gpurun_out/s2_ride_base8.json 2> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --shape native --iters 100 --rounds 1 --env FO_BWD_RIDE=1 --env FO_BWD_RIDE=0 --env FO_BWD_RIDE=2 > gpurun_out/s2_ride_native8.json 2>> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --batch 1 --rounds 1 --env FO_BWD_RIDE=1 --env FO_BWD_RIDE=0 --env FO_BWD_RIDE=2 > gpurun_out/s2_ride_base1.json 2>> gpurun_out/s2_step_ab.err
cat gpurun_out/s2_ride_*.json; tail -3 gpurun_out/s2_step_ab.err
