# tuning sweep: shared-memory plan of the build / apply kernels (env hooks in kbbq_b200.cu: plan_smem)
for cfg in "4 32 3" "4 32 2" "3 32 2" "4 16 3" "3 16 3" "2 32 4"; do
  set -- $cfg
  KBBQ_KPS=$1 KBBQ_DREP=$2 KBBQ_MIN_STAGES=$3 timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_sweep_k$1_d$2_s$3.log 2>&1; echo "bench kps<=$1 drep<=$2 stages>=$3 rc=$?"
  KBBQ_KPS=$1 KBBQ_DREP=$2 KBBQ_MIN_STAGES=$3 python - <<PY
import json,sys
sys.path.insert(0,'kbbq-py_b200')
from kbbq import _native
d=json.loads(open('gpurun_out/bench_sweep_k$1_d$2_s$3.log').read().strip().splitlines()[-1])
pb,pa=_native.plan_info(150,1,6,3),_native.plan_info(150,1,6,2)
print('  build k%d s%d d%d | apply k%d s%d d%d |'%(pb['kps'],pb['stages'],pb['drep'],pa['kps'],pa['stages'],pa['drep']), {k:round(v['ms'],3) for k,v in d['kernels'].items()}, round(d['value']/1e9,1), 'Gbases/s')
PY
done
