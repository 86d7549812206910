# tools/gpu_final.sh TAG: GPU parity suite, smoke, the default bench line, the reference arm and the other BASELINE configs
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
T=$1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-160 gpurun_out/${T}_bench.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; cut -c1-200 gpurun_out/${T}_bench_ref.json
for w in u3 u5 cyclic6 rollout; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; cut -c1-120 gpurun_out/${T}_bench_$w.json
done
