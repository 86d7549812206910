cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "overlap or prepared or shards" > gpurun_out/r2n_pytest.log 2>&1; tail -3 gpurun_out/r2n_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -2 gpurun_out/r2n_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2n_bench.json').read())
print(round(d['value']/1e6,1), 'M env-steps/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']/1e6,1))
print(json.dumps(d['cyclic6']))"
