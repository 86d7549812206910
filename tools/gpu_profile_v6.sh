set -x
cd $GRAFT_REPO_ROOT
for w in cyclic6 u3 u5 rollout; do timeout 300 python bench.py --workload $w --steps 3 --warmup 1 > gpurun_out/v6_bench_$w.json 2> gpurun_out/v6_bench_$w.err; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01_v6_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/v6_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_run -s 3 -c 1 -o gpurun_out/v6_krun python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/v6_n.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_run_wide -s 1 -c 1 -o gpurun_out/v6_kwide python bench.py --workload cyclic6 --episodes 256 --steps 1 --warmup 1 --no-cpu > gpurun_out/v6_nw.log 2>&1
ls -la gpurun_out
