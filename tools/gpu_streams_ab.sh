# tools/gpu_streams_ab.sh: stream-table size / occupancy A/B of the stream runner on cyclic-6
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in ":1184" "_k256c4:2368" "_k192c5:2960" "_k128c6:3552"; do
  lib=${v%%:*}; slots=${v#*:}
  for E in 1024 8192; do
    BBENV_LIB=$PWD/deepgroebner_b200/libbbenv$lib.so timeout 900 python bench.py --workload cyclic6 --episodes $E --slots $slots --steps 2 --warmup 1 --no-cpu > gpurun_out/sab${lib}_$E.json 2> gpurun_out/sab${lib}_$E.err || tail -3 gpurun_out/sab${lib}_$E.err
    python -c "
import json
d=json.loads(open('gpurun_out/sab${lib}_$E.json').read())
print('variant [$lib] episodes $E:', round(d['ms_per_step'],1), 'ms; adds/s', round(d['additions_per_sec']/1e6,1), 'M; slots', d['config']['slots'])"
  done
done
