timeout 600 python -m pytest tests/test_gpu_train.py -q -k "unit or train_mode_forward" 2>&1 | grep -E "^E   |passed|failed" | cut -c1-300 | head -12
for sp in 0 1; do echo "split $sp"; GFR_BN_BWD_SPLIT=$sp timeout 200 python tools/time_bn.py | head -3; done
for sp in 0 1 0 1; do
GFR_BN_BWD_SPLIT=$sp timeout 300 python bench.py --workload train --no-gpu-ref --cpu-faces 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
t=d.get('train', d)
print('split $sp train', t['value'], t['ms_per_step'], t['gpu_launches'])"
done
