set -x
mkdir -p gpurun_out
CHUNKS=2,4,8 python profiles/e2e_probe.py > gpurun_out/s2_e2e_probe.txt 2>&1
cat gpurun_out/s2_e2e_probe.txt | tail -12
