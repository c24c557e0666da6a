# tuning sweep: shared-memory plan of the build / apply kernels (env hooks in kbbq_b200.cu: plan_smem)
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_sweep.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_sweep.log
for cfg in "4 32" "4 16" "3 16" "2 32" "2 16"; do
  set -- $cfg
  KBBQ_KPS=$1 KBBQ_DREP=$2 timeout 150 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_sweep_k$1_d$2.log 2>&1; echo "bench kps<=$1 drep<=$2 rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_sweep_k$1_d$2.log').read().strip().splitlines()[-1])
print({k:round(v['ms'],3) for k,v in d['kernels'].items()}, round(d['value']/1e9,1), 'Gbases/s')
PY
done
