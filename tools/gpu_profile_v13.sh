set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/v13_pytest.log 2>&1; tail -3 gpurun_out/v13_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/v13_bench.json 2> gpurun_out/v13_bench.err; cut -c1-300 gpurun_out/v13_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01_v13_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/v13_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_run -s 3 -c 1 -o gpurun_out/v13_krun python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/v13_n.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_prepare -s 3 -c 1 -o gpurun_out/v13_kprep python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/v13_np.log 2>&1
ls -la gpurun_out
