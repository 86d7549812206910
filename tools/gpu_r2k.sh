cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "overlap or prepared" > gpurun_out/r2k_pytest.log 2>&1; tail -3 gpurun_out/r2k_pytest.log
bash tools/gpu_ab.sh stock:--no-pipeline
