cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "round2 or lead_monomials_env or golden or batched_step" > gpurun_out/r2j_pytest.log 2>&1; tail -3 gpurun_out/r2j_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -2 gpurun_out/r2j_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2j_bench.json').read())
print(round(d['value']/1e6,1), 'M env-steps/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']/1e6,1), d['kernel_ms'], d['step_ms'], 'dropin', d['dropin_n1']['value'], d['dropin_n1']['reference_cython'].get('value'))"
timeout 300 python bench.py --steps 10 --warmup 3 --episodes 65536 --no-cpu --no-extras > gpurun_out/r2j_bench64k.json 2> gpurun_out/r2j_bench64k.err
python -c "
import json
d=json.loads(open('gpurun_out/r2j_bench64k.json').read())
print('65536 episodes:', round(d['value']/1e6,1), 'M env-steps/s', round(d['ms_per_step'],4), 'ms', d['kernel_ms'])"
bash tools/gpu_profile.sh r2j_krun k_run --no-pipeline
