cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "lead_monomials or batched_step or golden or cyclic6_observation or server or compaction or rollout or policy or smoke or max_episode" > gpurun_out/r2r_pytest.log 2>&1; tail -3 gpurun_out/r2r_pytest.log
python -c "
import sys, json
sys.path.insert(0, '.')
import torch, bench
bench.DIST='3-20-10-weighted'; bench.STRATEGY='degree'
dev=torch.device('cuda',0); flush=torch.empty(160<<20,dtype=torch.uint8,device=dev)
print(json.dumps(bench.extra_step_api(torch, dev, 0, flush)))
print(json.dumps(bench.extra_dropin_n1(2.0)))"
timeout 600 python bench.py --workload rollout --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rollout', round(d['value']/1e6,1), 'M env-steps/s')"
