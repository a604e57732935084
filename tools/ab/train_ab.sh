timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_trajectory.py tests/test_gpu_train_loop.py tests/test_gpu_train_lighting_transfer.py tests/test_gpu_optimizer_and_caches.py -q -x 2>&1 | tail -5
for cfg in "0 0" "1 0" "0 1" "1 1" "1 1"; do set -- $cfg
GFR_TRAIN_DEDUP_D=$1 GFR_TRAIN_PACK_PLAN=$2 timeout 300 python bench.py --workload train --no-gpu-ref --cpu-faces 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
t=d.get('train', d)
print('dedup $1 plan $2 train', t['value'], t['ms_per_step'], t['gpu_launches'])"
done
