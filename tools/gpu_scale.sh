# tools/gpu_scale.sh "<N ...>" "<workload ...>": the driver's N-GPU launch of bench.py (our arm) for every N and workload
# (episodes = configs[1] weak scaling; u3 / u5 = configs[2], 65536 episodes in total sharded over the ranks), plus the 2-rank
# NCCL all-gather test when the box has at least two GPUs.  Lines go to gpurun_out/scale_<workload>_<N>.json.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for w in ${2:-episodes}; do
  for N in ${1:-8}; do
    if [ $N = 1 ]; then launch="python"; else launch="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"; fi
    BB_BENCH_WORKLOAD=$w timeout 900 $launch bench.py --gpus $N --steps 20 --warmup 3 --no-extras > gpurun_out/scale_${w}_$N.json 2> gpurun_out/scale_${w}_$N.err
    python -c "
import json
d=json.loads(open('gpurun_out/scale_${w}_$N.json').read())
print('$w', $N, 'GPUs', round(d['value']/1e6,1), 'M env-steps/s', round(d['ms_per_step'],4), 'ms', d.get('per_rank'), 'e2e', round(d['e2e']['value']/1e6,1), 'parity', d['parity']['episodes_checked'], d['parity']['mismatches'])" || tail -5 gpurun_out/scale_${w}_$N.err
  done
done
if [ "$(python -c 'import torch; print(torch.cuda.device_count())')" -ge 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_nccl.py -m gpu -x -q > gpurun_out/scale_nccl_pytest.log 2>&1; tail -2 gpurun_out/scale_nccl_pytest.log
fi
