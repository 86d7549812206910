cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; tail -5 gpurun_out/r2o_pytest.log
bash tools/gpu_rep.sh 2
python tools/exp_cyclic_chain.py 1024
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; tail -2 gpurun_out/r2o_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2o_bench.json').read())
print(round(d['value']/1e6,1), 'M env-steps/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']/1e6,1))
print(json.dumps(d['dropin_n1'])); print(json.dumps(d['cyclic6']))"
