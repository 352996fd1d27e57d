set -x
mkdir -p gpurun_out
python profiles/step_ab.py --env FO_PDL=0,FO_RANK_FAST=0 --env FO_PDL=31,FO_RANK_FAST=0 --env FO_PDL=31,FO_RANK_FAST=1 > gpurun_out/s2_step_ab_base8.json 2> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --batch 1 --env FO_PDL=0,FO_RANK_FAST=0 --env FO_PDL=31,FO_RANK_FAST=0 --env FO_PDL=31,FO_RANK_FAST=1 > gpurun_out/s2_step_ab_base1.json 2>> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --shape native --iters 100 --env FO_PDL=0,FO_RANK_FAST=0 --env FO_PDL=31,FO_RANK_FAST=0 --env FO_PDL=31,FO_RANK_FAST=1 > gpurun_out/s2_step_ab_native8.json 2>> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --shape stress --batch 2 --iters 100 --env FO_PDL=0,FO_RANK_FAST=0 --env FO_PDL=31,FO_RANK_FAST=0 --env FO_PDL=31,FO_RANK_FAST=1 > gpurun_out/s2_step_ab_stress2.json 2>> gpurun_out/s2_step_ab.err
cat gpurun_out/s2_step_ab_*.json; tail -5 gpurun_out/s2_step_ab.err
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/s2_pytest2.log
