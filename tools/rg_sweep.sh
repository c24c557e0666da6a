# tuning sweep for the several-read-groups path (env hooks in kbbq_b200.cu: plan_kernel / plan_smem)
R=${1:-8}; L=${2:-150}; N=${3:-10000000}
for np in 4 8; do for dr in 16 32; do for k in 1 2 3; do for ms in 3 8; do
  export KBBQ_NPROD=$np KBBQ_DREP=$dr KBBQ_KPS=$k KBBQ_MIN_STAGES=$ms
  timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --read-groups $R --read-len $L --reads $N > /tmp/rg_sweep.log 2>&1
  python - <<PY
import json,sys
sys.path.insert(0,'kbbq-py_b200')
from kbbq import _native
try:
    d=json.loads(open('/tmp/rg_sweep.log').read().strip().splitlines()[-1])
    pb,pa=_native.plan_info($L,$R,6,3),_native.plan_info($L,$R,6,2)
    f=lambda p:'np%d ng%d k%d s%d d%d'%(p['nprod'],p['ng'],p['kps'],p['stages'],p['drep'])
    print('R=$R L=$L req np$np d$dr k$k ms$ms | build',f(pb),'| apply',f(pa),'|',{k:round(v['ms'],3) for k,v in d['kernels'].items()}, round(d['value']/1e9,1),'Gbases/s')
except Exception as e:
    print('R=$R L=$L req np$np d$dr k$k ms$ms FAILED', e)
PY
done; done; done; done
