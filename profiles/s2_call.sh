set -x
mkdir -p gpurun_out
python profiles/step_ab.py --batch 1 --env FO_BWD_RIDE=1 --env FO_BWD_RIDE=2 > gpurun_out/s2_m6_base1.json 2> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --batch 2 --rounds 1 --env FO_BWD_RIDE=1 --env FO_BWD_RIDE=2 > gpurun_out/s2_m6_base2.json 2>> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --shape stress --batch 2 --iters 100 --rounds 1 --env FO_BWD_RIDE=1 --env FO_BWD_RIDE=2  --env FO_BWD_RIDE=0 > gpurun_out/s2_m6_stress2.json 2>> gpurun_out/s2_step_ab.err
cat gpurun_out/s2_m6_base1.json gpurun_out/s2_m6_base2.json gpurun_out/s2_m6_stress2.json; tail -3 gpurun_out/s2_step_ab.err
