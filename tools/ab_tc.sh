python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/pytest_gpu.log
for k in ${KERNELS:-tc1 tcp tc}; do
  for w in ${WORKLOADS:-cfg5 cfg2}; do
    DMK_FD_KERNEL=$k timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-others --e2e-steps 0 > gpurun_out/ab_${k}_${w}.json 2> gpurun_out/ab_${k}_${w}.err; 
    python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/ab_${k}_${w}.json").read().strip().splitlines()[-1])
    print("$k $w", j["ms_per_step"], j["roofline"]["frac"], j["roofline"]["kernel"])
except Exception as e:
    print("$k $w failed", e)
PY
  done
done
