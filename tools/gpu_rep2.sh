cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in b2s3 b3s3 b2s4 stock; do
  if [ $v = stock ]; then unset BBENV_LIB; else export BBENV_LIB=$GRAFT_REPO_ROOT/deepgroebner_b200/libbbenv_$v.so; fi
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/rep_$v.json 2> gpurun_out/rep_$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/rep_$v.json').read())
print('$v', round(d['value']/1e6,1), 'M env-steps/s', round(d['ms_per_step'],4), 'ms; max step', max(d['step_ms']), 'steps', d['step_ms'][:6])"
done
