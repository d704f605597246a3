W=${1:-cfg5}; U=${2:-20000}; PAT=${3:-fd_ws}
python bench.py --workload $W --users $U --steps 2 --warmup 1 --no-cpu --no-others --e2e-steps 0 > gpurun_out/plain_$W.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:$PAT -s 2 -c 1 -o gpurun_out/prof_${PAT}_$W -f python bench.py --workload $W --users $U --steps 2 --warmup 1 --no-cpu --no-others --e2e-steps 0 > gpurun_out/ncu_$W.log 2>&1
tail -2 gpurun_out/ncu_$W.log
