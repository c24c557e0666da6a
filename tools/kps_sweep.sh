set -x
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu7.log
for k in 1 2 3 4; do
  KBBQ_KPS=$k timeout 150 python bench.py --steps 5 --warmup 3 > gpurun_out/bench7_k$k.log 2>&1; echo "bench k=$k rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/bench7_k$k.log').read().strip().splitlines()[-1]);print({k:round(v['ms'],3) for k,v in d['kernels'].items()},d['value'])"
done
