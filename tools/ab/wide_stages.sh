for lib in "" _w3 _w4 _w5; do
  echo "== lib$lib"
  GFR_LIB_PATH=/root/repo/geomconsistentfr_b200/csrc/libgfr_b200$lib.so timeout 300 python tools/time_train_convs.py 2>&1 | tail -13
  GFR_LIB_PATH=/root/repo/geomconsistentfr_b200/csrc/libgfr_b200$lib.so timeout 300 python bench.py --workload train --no-gpu-ref --cpu-faces 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
t=d.get('train', d)
print('train', t['value'], t['ms_per_step'])"
done
