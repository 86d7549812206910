cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv
$TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/multi_${N}_episodes.json 2> gpurun_out/multi_${N}_episodes.err
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/multi_1_episodes.json 2>/dev/null
CUDA_VISIBLE_DEVICES=1 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/multi_1b_episodes.json 2>/dev/null
for f in gpurun_out/multi_*episodes.json; do echo "== $f"; python -c "
import json,sys
d=json.loads(open('$f').read())
print(d['value'], d['ms_per_step'], d.get('per_rank'), d['clocks'], d['e2e'])"; done
tail -3 gpurun_out/multi_${N}_episodes.err
