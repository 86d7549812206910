cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in "$@" stock; do
  if [ $v = stock ]; then unset BBENV_LIB; else export BBENV_LIB=$GRAFT_REPO_ROOT/deepgroebner_b200/libbbenv_$v.so; fi
  timeout 300 python bench.py --workload rollout --steps 10 --warmup 3 --no-cpu > gpurun_out/abr_$v.json 2> gpurun_out/abr_$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/abr_$v.json').read())
print('$v', round(d['value']/1e6,2), 'M env-steps/s', round(d['ms_per_step'],2), 'ms')"
done
