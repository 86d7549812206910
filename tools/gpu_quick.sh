# tools/gpu_quick.sh: GPU parity tests + one bench line (the loop of every kernel change)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; tail -3 gpurun_out/q_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; tail -2 gpurun_out/q_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/q_bench.json').read())
print(round(d['value']/1e6,1), 'M env-steps/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']/1e6,1))"
