# tools/gpu_ncu_wide.sh TAG: ncu --set full capture of k_run_wide on 256 cyclic-6 episodes (the launch lasts as long as its longest episode)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_run_wide -s 1 -c 1 -o gpurun_out/$1_kwide python bench.py --workload cyclic6 --episodes 256 --steps 1 --warmup 1 --no-cpu > gpurun_out/$1_nw.log 2>&1
ls -la gpurun_out/$1_kwide.ncu-rep; tail -3 gpurun_out/$1_nw.log
