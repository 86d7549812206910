# tools/gpu_wide_iter2.sh <tag>: gpu_wide_iter.sh + the cyclic-6 bench lines at 1024 / 8192 episodes (296 CTA slots)
cd $GRAFT_REPO_ROOT
bash tools/gpu_wide_iter.sh $1
SLOTS=296 bash tools/gpu_cyc_bench.sh
