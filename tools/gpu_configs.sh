# tools/gpu_configs.sh <tag>: one bench line per BASELINE config on one box (configs[1] is gpu_base.sh's)
tag=${1:-cfg}
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for w in rollout cyclic6; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/${tag}_$w.json 2> gpurun_out/${tag}_$w.err
  python -c "
import json
d=json.loads(open('gpurun_out/${tag}_$w.json').read())
print('$w', round(d['value']/1e6,3), 'M env-steps/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']/1e6,3), 'adds/s', round(d.get('additions_per_sec',0)/1e6,1), d.get('parity'), d.get('cpu_baseline',{}).get('value'))" || tail -5 gpurun_out/${tag}_$w.err
done
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "server" > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
python -c "
import sys, time, numpy as np
sys.path.insert(0, '.')
import bench
bench.DIST='3-20-10-weighted'
print(bench.extra_dropin_n1(2.0))"
