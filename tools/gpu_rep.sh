# tools/gpu_rep.sh [n] [bench args]: the bench line n times on one box (run-to-run spread, one-off stalls)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
n=${1:-3}; shift
for i in $(seq 1 $n); do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras "$@" > gpurun_out/rep_$i.json 2> gpurun_out/rep_$i.err
  python -c "
import json
d=json.loads(open('gpurun_out/rep_$i.json').read())
print($i, round(d['value']/1e6,1), 'M env-steps/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']/1e6,1), 'max step', max(d['step_ms']), 'steps', d['step_ms'])"
done
