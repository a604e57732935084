export GFR_LIB_PATH=/root/repo/geomconsistentfr_b200/csrc/libgfr_b200_g4.so
timeout 100 python bench.py --workload train --no-gpu-ref --cpu-faces 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
t=d.get('train', d)
print('lib_g4 train', t['value'], t['ms_per_step'])"
timeout 60 python tools/time_train_convs.py 2>&1 | tail -10 | head -6
timeout 200 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_trajectory.py -q -x 2>&1 | grep -E "^E   |passed|failed" | cut -c1-200 | head -4
