# tools/gpu_ab.sh [-a "bench args"] V1 V2 ...: bench.py against libbbenv_<V>.so variants (csrc/Makefile `variant`), then the
# stock library; a variant name of the form  stock:<args>  runs the stock library with extra bench.py arguments
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
extra=""
if [ "$1" = "-a" ]; then extra="$2"; shift 2; fi
for v in "$@" stock; do
  args="$extra"; name=$v
  case $v in stock:*) args="$extra ${v#stock:}"; name=stock_$(echo "${v#stock:}" | tr -c 'a-zA-Z0-9\n' '_'); unset BBENV_LIB;;
             stock) unset BBENV_LIB;;
             *) export BBENV_LIB=$GRAFT_REPO_ROOT/deepgroebner_b200/libbbenv_$v.so;; esac
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras $args > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python -c "
import json
d=json.loads(open('gpurun_out/ab_$name.json').read())
print('$name', round(d['value']/1e6,1), 'M env-steps/s', round(d['ms_per_step'],4), 'ms', 'e2e', round(d['e2e']['value']/1e6,1), 'kernels', {k: round(v,4) for k,v in d['kernel_ms'].items() if k!='how'}, 'steps', d.get('step_ms'))" || tail -3 gpurun_out/ab_$name.err
done
