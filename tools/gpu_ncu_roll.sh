cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 2 -c 1 -o gpurun_out/$1_kroll python bench.py --workload rollout --steps 2 --warmup 1 --no-cpu > gpurun_out/$1_nr.log 2>&1
tail -2 gpurun_out/$1_nr.log; ls -la gpurun_out/$1_kroll.ncu-rep
