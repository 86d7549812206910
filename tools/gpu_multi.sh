# tools/gpu_multi.sh N: the driver's multi-GPU launch of bench.py on N GPUs of one box (both arms + the sharded config)
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/multi_${N}_episodes.json 2> gpurun_out/multi_${N}_episodes.err
$TR --master-port 29512 bench.py --gpus $N --workload u3 --steps 5 --warmup 3 > gpurun_out/multi_${N}_u3.json 2> gpurun_out/multi_${N}_u3.err
$TR --master-port 29513 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/multi_${N}_ref.json 2> gpurun_out/multi_${N}_ref.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/multi_1_ref.json 2> gpurun_out/multi_1_ref.err
for f in gpurun_out/multi_*.json; do echo "== $f"; cut -c1-400 $f; done
tail -3 gpurun_out/multi_${N}_episodes.err
