# scratch: Morton key width sweep (bits per axis) for the target tree and the source ordering
for cfg in "13 12" "13 13" "13 14" "13 15" "12 12" "14 13"; do set -- $cfg
WAVECU_TGT_BITS=$1 WAVECU_SRC_BITS=$2 timeout 300 python bench.py --skip-cpu --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tgt $1 src $2: ms/step %.3f e2e %.3f launch_us %.1f'%(d['ms_per_step'],d['e2e']['ms_per_step'],1e3*d['roofline']['mean_launch_ms']), d['breakdown_ms_per_step'])"
done
