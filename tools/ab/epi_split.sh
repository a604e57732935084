timeout 600 python -m pytest tests/test_gpu_conv_p16.py tests/test_gpu_relight.py tests/test_gpu_range_safety.py tests/test_gpu_lighting_transfer.py -q 2>&1 | grep -E "^E   |passed|failed" | cut -c1-300 | head -12
for m in 0 1; do echo "split $m"; GFR_P16_EPI_SPLIT=$m timeout 200 python tools/time_conv_p16.py 2>&1 | head -2; done
for m in 0 1 0 1; do GFR_P16_EPI_SPLIT=$m timeout 300 python bench.py --workload forward --no-gpu-ref --cpu-faces 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('split', $m, d['value'], d['e2e']['value'], d['latency']['ms_per_step'], d['roofline_cnn']['ms_per_launch'])"; done
