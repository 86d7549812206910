cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; tail -3 gpurun_out/r2l_pytest.log
bash tools/gpu_ab.sh stock:--no-pipeline
python tools/exp_pipeline.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_prepare_lanes -s 2 -c 1 -o gpurun_out/r2l_kprep python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --no-pipeline > gpurun_out/r2l_kprep_ncu.log 2>&1; ls -la gpurun_out/r2l_kprep.ncu-rep
