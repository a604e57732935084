timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_trajectory.py -q -x 2>&1 | tail -3
timeout 200 python tools/time_bn.py
timeout 300 python bench.py --workload train --no-gpu-ref --cpu-faces 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
t=d.get('train', d)
print('train', t['value'], t['ms_per_step'], t['gpu_launches'])"
