timeout 600 python -m pytest tests/test_gpu_dp_nccl.py -q 2>&1 | grep -E "^E   |passed|failed|skipped" | cut -c1-300 | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-gpu-ref --cpu-faces 1 > gpurun_out/r02_bench_w_n2.json 2> gpurun_out/r02_bench_w_n2.err
tail -c 300 gpurun_out/r02_bench_w_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_w_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['train']['value'], d['train']['ms_per_step'], d['train']['allreduce'], d['sweep']['value'])"
