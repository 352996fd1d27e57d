set -x
mkdir -p gpurun_out
python profiles/step_ab.py --rounds 1 --env FO_PDL=31 > gpurun_out/s2_nodg_ref.json 2> gpurun_out/s2_step_ab.err
FUSIONOCC_B200_LIB=fusionocc_b200/lib/libfusionocc_b200_nodg.so python profiles/step_ab.py --rounds 1 --env FO_PDL=31 > gpurun_out/s2_nodg_a.json 2>> gpurun_out/s2_step_ab.err
FUSIONOCC_B200_LIB=fusionocc_b200/lib/libfusionocc_b200_nodg6.so python profiles/step_ab.py --rounds 1 --env FO_PDL=31 > gpurun_out/s2_nodg_b.json 2>> gpurun_out/s2_step_ab.err
cat gpurun_out/s2_nodg_*.json
