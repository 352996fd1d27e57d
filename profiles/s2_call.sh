set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err
tail -c 400 gpurun_out/r02b_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r02b_bench_n2.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','n_gpus','e2e','gather','c3_inference'): print(k, json.dumps(d.get(k))[:500])
"
