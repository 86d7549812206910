# tools/gpu_ab_cyc.sh V1 V2 ...: cyclic-6 bench line against libbbenv_<V>.so variants, then the stock library
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in "$@" stock; do
  if [ $v = stock ]; then unset BBENV_LIB; else export BBENV_LIB=$GRAFT_REPO_ROOT/deepgroebner_b200/libbbenv_$v.so; fi
  timeout 300 python bench.py --workload cyclic6 --steps 2 --warmup 3 --no-cpu > gpurun_out/abc_$v.json 2> gpurun_out/abc_$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/abc_$v.json').read())
print('$v', round(d['value']/1e6,4), 'M env-steps/s', round(d['ms_per_step'],1), 'ms', 'slots', d['config']['slots'])"
done
