# tools/gpu_scale.sh N: the driver's N-GPU launch of bench.py (our arm only) + per-rank times
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
python -c "
import json
d=json.loads(open('gpurun_out/scale_$N.json').read())
print($N, 'GPUs', round(d['value']/1e6,1), 'M env-steps/s', d['ms_per_step'], d.get('per_rank'), 'e2e', round(d['e2e']['value']/1e6,1))"
tail -2 gpurun_out/scale_$N.err
