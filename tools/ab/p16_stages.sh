for st in 2 3 4; do
echo "stages $st"; GFR_P16_STAGES=$st timeout 200 python tools/time_conv_p16.py 2>&1 | head -2
GFR_P16_STAGES=$st timeout 300 python bench.py --workload forward --no-gpu-ref --cpu-faces 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stages', $st, d['value'], d['e2e']['value'], d['latency']['ms_per_step'], d['roofline_cnn']['ms_per_launch'])"; done
