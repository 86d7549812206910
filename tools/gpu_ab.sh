# tools/gpu_ab.sh V1 V2 ...: bench.py against libbbenv_<V>.so variants (csrc/Makefile `variant`), then the stock library
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in "$@" stock; do
  if [ $v = stock ]; then unset BBENV_LIB; else export BBENV_LIB=$GRAFT_REPO_ROOT/deepgroebner_b200/libbbenv_$v.so; fi
  timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/ab_$v.json').read())
print('$v', round(d['value']/1e6,1), 'M env-steps/s', round(d['ms_per_step'],4), 'ms', 'slots', d['config']['slots'])"
done
