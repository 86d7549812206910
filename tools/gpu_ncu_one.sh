cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --clock-control none --import-source on -k regex:k_run_wide -s 1 -c 1 -o gpurun_out/$1_one python tools/exp_cyclic_one.py > gpurun_out/$1_one.log 2>&1
tail -3 gpurun_out/$1_one.log; ls -la gpurun_out/$1_one.ncu-rep
