# tools/gpu_cyc_bench.sh: cyclic-6 bench lines (BASELINE configs[4]) at 1024 and 8192 episodes per launch
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for E in 1024 8192; do
  timeout 900 python bench.py --workload cyclic6 --episodes $E --slots ${SLOTS:-1184} --steps 2 --warmup 1 --no-cpu > gpurun_out/cyc_$E.json 2> gpurun_out/cyc_$E.err || tail -3 gpurun_out/cyc_$E.err
  python -c "
import json
d=json.loads(open('gpurun_out/cyc_$E.json').read())
print('episodes $E:', round(d['ms_per_step'],1), 'ms; adds/s', round(d['additions_per_sec']/1e6,1), 'M; env-steps/s', round(d['value']/1e6,3), 'M; slots', d['config']['slots'])"
done
