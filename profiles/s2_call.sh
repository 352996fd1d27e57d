set -x
mkdir -p gpurun_out
python profiles/step_ab.py --env FO_BWD_HALF=0 --env FO_BWD_HALF=1 > gpurun_out/s2_emit_base8.json 2> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --shape native --iters 100 --rounds 1 --env FO_BWD_HALF=0 --env FO_BWD_HALF=1 > gpurun_out/s2_emit_native8.json 2>> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --batch 1 --rounds 1 --env FO_BWD_HALF=0 --env FO_BWD_HALF=1 > gpurun_out/s2_emit_base1.json 2>> gpurun_out/s2_step_ab.err
cat gpurun_out/s2_emit_base8.json gpurun_out/s2_emit_native8.json gpurun_out/s2_emit_base1.json; tail -3 gpurun_out/s2_step_ab.err
timeout 600 python -m pytest tests/test_gpu_switches.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
