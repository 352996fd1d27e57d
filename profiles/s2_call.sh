set -x
mkdir -p gpurun_out
python profiles/step_ab.py --env FO_BWD_PIX=2 --env FO_BWD_PIX=4 > gpurun_out/s2_pixdb_base8.json 2> gpurun_out/s2_step_ab.err
FUSIONOCC_B200_LIB=fusionocc_b200/lib/libfusionocc_b200_t128.so python profiles/step_ab.py --rounds 1 --env FO_BWD_PIX=4 > gpurun_out/s2_pixdb_t128_base8.json 2>> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --batch 1 --rounds 1 --env FO_BWD_PIX=2 --env FO_BWD_PIX=4 > gpurun_out/s2_pixdb_base1.json 2>> gpurun_out/s2_step_ab.err
python profiles/step_ab.py --shape native --iters 100 --rounds 1 --env FO_BWD_PIX=2 --env FO_BWD_PIX=4 > gpurun_out/s2_pixdb_native8.json 2>> gpurun_out/s2_step_ab.err
cat gpurun_out/s2_pixdb_*.json; tail -3 gpurun_out/s2_step_ab.err
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/s2_pytest2.log
