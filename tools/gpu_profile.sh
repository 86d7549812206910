# tools/gpu_profile.sh <tag> <kernel regex> [bench.py arguments]: one `ncu --set full` capture of the named kernel of a
# bench.py run (the recipe of /opt/skills/guides/B200_PROFILING.md), brought back as gpurun_out/<tag>.ncu-rep; summarise it
# here with tools/ncu_summary.sh gpurun_out/<tag>.ncu-rep profiles/<prefix>.  A number printed under ncu is never a bench value.
tag=$1; kre=$2; shift 2
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kre -s 3 -c 1 -o gpurun_out/$tag \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-extras "$@" > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log; ls -la gpurun_out/$tag.ncu-rep
