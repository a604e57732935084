timeout 600 python -m pytest tests/test_gpu_train.py -q 2>&1 | grep -E "^E   |passed|failed" | cut -c1-300 | head -12
timeout 200 python tools/time_bn.py | tail -1
for a in 1 1; do
timeout 300 python bench.py --workload train --no-gpu-ref --cpu-faces 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
t=d.get('train', d)
print('train', t['value'], t['ms_per_step'], t['gpu_launches'])"
done
