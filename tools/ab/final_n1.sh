timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E   |passed|failed" | cut -c1-300 | head -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02_bench_final3_n1.json 2> gpurun_out/r02_bench_final3_n1.err; tail -c 200 gpurun_out/r02_bench_final3_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_final3_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['latency']['ms_per_step'], d['gpu_launches'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['roofline']['issue']['frac'])
print(d['train']['value'], d['train']['ms_per_step'], d['sweep']['value'], d['roofline_cnn']['ms_per_launch'], d['roofline_cnn']['frac'])
print(d['cpu_baseline']['value'], d['reference_gpu']['value'], d['clocks'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
