# tools/gpu_wide_iter.sh <tag>: one iteration on the long-polynomial runner: parity tests of the CTA-per-environment
# runner, chain timing of the longest cyclic-6 episodes, one ncu source-counter capture of a mid-length episode alone
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cta_per_environment or cyclic6_seeded or fixed_ideals or preparation" > gpurun_out/$1_pytest.log 2>&1; tail -3 gpurun_out/$1_pytest.log
timeout 600 python tools/exp_cyclic_chain.py 1024 > gpurun_out/$1_chain.log 2>&1; cat gpurun_out/$1_chain.log
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --clock-control none --import-source on -k regex:k_run_wide -s 1 -c 1 -o gpurun_out/$1_one python tools/exp_cyclic_one.py 1237 > gpurun_out/$1_one.log 2>&1
tail -3 gpurun_out/$1_one.log; ls -la gpurun_out/$1_one.ncu-rep
