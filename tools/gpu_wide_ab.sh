# tools/gpu_wide_ab.sh: CTA shape A/B of the long-polynomial runner on cyclic-6 (1024 and 8192 episodes per launch)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in "" _w4c4 _w4c6 _w8c3; do
  for E in 1024 8192; do
    BBENV_LIB=$PWD/deepgroebner_b200/libbbenv$v.so timeout 600 python bench.py --workload cyclic6 --episodes $E --steps 2 --warmup 1 --no-cpu > gpurun_out/ab${v}_$E.json 2> gpurun_out/ab${v}_$E.err || tail -3 gpurun_out/ab${v}_$E.err
    python -c "
import json,sys
d=json.loads(open('gpurun_out/ab${v}_$E.json').read())
print('variant [$v] episodes $E:', round(d['ms_per_step'],1), 'ms; adds/s', round(d['additions_per_sec']/1e6,1), 'M; slots', d['config']['slots'])"
  done
done
