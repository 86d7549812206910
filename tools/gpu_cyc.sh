# tools/gpu_cyc.sh: the long-polynomial runner (k_run_wide): its parity tests and the cyclic-6 bench line
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "wide or cyclic" > gpurun_out/c_pytest.log 2>&1; tail -3 gpurun_out/c_pytest.log
timeout 600 python bench.py --workload cyclic6 --steps 3 --warmup 3 --no-cpu > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; tail -2 gpurun_out/c_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/c_bench.json').read())
print(round(d['value']/1e6,4), 'M env-steps/s', round(d['ms_per_step'],1), 'ms; adds/s', round(d['additions_per_sec']/1e6,1), d['config']['workload'])"
