# tools/gpu_pol.sh: the fused policy head / rollout: its tests and the rollout bench line (BASELINE configs[3])
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_policy_rollout.py -m gpu -x -q > gpurun_out/p_pytest.log 2>&1; tail -15 gpurun_out/p_pytest.log
timeout 600 python bench.py --workload rollout --steps 10 --warmup 3 --no-cpu > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; tail -2 gpurun_out/p_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/p_bench.json').read())
print(round(d['value']/1e6,2), 'M env-steps/s', round(d['ms_per_step'],2), 'ms; e2e', round(d['e2e']['value']/1e6,2), d['config']['workload'][:80])"
