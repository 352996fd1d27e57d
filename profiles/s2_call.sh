set -x
mkdir -p gpurun_out
python profiles/step_ab.py --env FO_RANK_SLOTFREE=0 --env FO_RANK_SLOTFREE=1 > gpurun_out/s2_slotfree_base8.json 2> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --shape native --iters 100 --rounds 1 --env FO_RANK_SLOTFREE=0 --env FO_RANK_SLOTFREE=1 > gpurun_out/s2_slotfree_native8.json 2>> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --batch 1 --rounds 1 --env FO_RANK_SLOTFREE=0 --env FO_RANK_SLOTFREE=1 > gpurun_out/s2_slotfree_base1.json 2>> gpurun_out/s2_step_ab.err
cat gpurun_out/s2_slotfree_*.json; tail -3 gpurun_out/s2_step_ab.err
timeout 900 python -m pytest tests/test_gpu_switches.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/s2_pytest3.log
