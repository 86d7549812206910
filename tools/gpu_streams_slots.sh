# tools/gpu_streams_slots.sh: concurrent-episode sweep of the stream runner on cyclic-6 (8192 episodes per launch)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for slots in 148 296 444 592 888; do
  timeout 900 python bench.py --workload cyclic6 --episodes 8192 --slots $slots --steps 2 --warmup 1 --no-cpu > gpurun_out/ss_$slots.json 2> gpurun_out/ss_$slots.err || tail -3 gpurun_out/ss_$slots.err
  python -c "
import json
d=json.loads(open('gpurun_out/ss_$slots.json').read())
print('slots $slots:', round(d['ms_per_step'],1), 'ms; adds/s', round(d['additions_per_sec']/1e6,1), 'M')"
done
