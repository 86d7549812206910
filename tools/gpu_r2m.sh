cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
bash tools/gpu_base.sh r2m
python -c "
import json
d=json.loads(open('gpurun_out/r2m_bench.json').read())
for k in ('kernel_ms','step_api','dropin_n1','cyclic6','with_gb','parity'): print(k, json.dumps(d.get(k))[:900])"
timeout 900 bash tools/gpu_profile.sh r2m_kwide k_run_wide --workload cyclic6 --episodes 148 --no-pipeline
