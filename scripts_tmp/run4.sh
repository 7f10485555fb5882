timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests2.log 2>&1; tail -4 gpurun_out/final_tests2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_2gpu.json') if l.startswith('{')][-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'], d['kernel_ms_per_step'])"
