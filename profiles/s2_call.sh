set -x
mkdir -p gpurun_out
python profiles/step_ab.py --env FO_BWD_HALF=0 --env FO_BWD_HALF=1 > gpurun_out/s2_mask_base8.json 2> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --batch 1 --rounds 1 --env FO_BWD_HALF=0 --env FO_BWD_HALF=1 > gpurun_out/s2_mask_base1.json 2>> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --shape native --iters 100 --rounds 1 --env FO_BWD_HALF=0 --env FO_BWD_HALF=1 > gpurun_out/s2_mask_native8.json 2>> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --shape stress --batch 2 --iters 100 --rounds 1 --env FO_BWD_HALF=0 --env FO_BWD_HALF=1 > gpurun_out/s2_mask_stress2.json 2>> gpurun_out/s2_step_ab.err
cat gpurun_out/s2_mask_*.json; tail -3 gpurun_out/s2_step_ab.err
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/s2_pytest2.log
