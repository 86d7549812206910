# tools/gpu_ncu_krun.sh TAG: one ncu --set full capture of k_run (third launch) of the default bench workload
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_run -s 3 -c 1 -o gpurun_out/$1_krun python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/$1_n.log 2>&1
ls -la gpurun_out/$1_krun.ncu-rep
