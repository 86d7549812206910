# tools/gpu_final.sh <tag>: what the driver runs at round end, in its order: GPU tests, smoke(), the reference arm, our arm
tag=${1:-final}
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; cut -c1-400 gpurun_out/${tag}_bench_ref.json
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read()); r=json.loads(open('gpurun_out/${tag}_bench_ref.json').read())
print(round(d['value']/1e6,1), 'M env-steps/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']/1e6,1), '; reference', round(r['value']/1e6,2), 'M; ratio', round(d['value']/r['value'],1), 'e2e ratio', round(d['e2e']['value']/r['e2e']['value'],1), 'same workload string', d['config']['workload']==r['config']['workload'])
for k in ('kernel_ms','step_api','dropin_n1','cyclic6','with_gb','parity','roofline'): print(k, json.dumps(d.get(k))[:700])"
