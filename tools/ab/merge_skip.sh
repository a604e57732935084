GFR_MERGE_SKIP=1 timeout 300 python -m pytest tests/test_gpu_relight.py tests/test_gpu_range_safety.py tests/test_gpu_lighting_transfer.py -q 2>&1 | tail -3
for m in 0 1 0 1; do GFR_MERGE_SKIP=$m timeout 300 python bench.py --workload forward --no-gpu-ref --cpu-faces 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('merge', $m, d['value'], d['e2e']['value'], d['latency']['ms_per_step'])"; done
