# tools/gpu_sanitize.sh: compute-sanitizer racecheck / memcheck over the block-level runner and the policy head (small cases)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cta_per_environment and cyclic-5" > gpurun_out/san_race_wide.log 2>&1; tail -6 gpurun_out/san_race_wide.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cta_per_environment and cyclic-5" > gpurun_out/san_mem_wide.log 2>&1; tail -4 gpurun_out/san_mem_wide.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_policy_rollout.py -m gpu -x -q -k "fused_rollout or (policy_head and 3-20-10-weighted-2-128)" > gpurun_out/san_mem_pol.log 2>&1; tail -4 gpurun_out/san_mem_pol.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_policy_rollout.py -m gpu -x -q -k "policy_head and 3-20-10-weighted-2-128" > gpurun_out/san_race_pol.log 2>&1; tail -4 gpurun_out/san_race_pol.log
