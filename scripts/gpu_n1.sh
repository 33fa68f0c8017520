#!/bin/bash
# single-GPU validation + profiling pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_n1.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_n1.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 400 --warmup 20 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?" >> gpurun_out/bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 27 --csv --log-file gpurun_out/swe_launches.csv python scripts/swe_steps.py 4096 4 > gpurun_out/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:swe_k -s 5 -c 5 -o gpurun_out/swe_kernels python scripts/swe_steps.py 4096 3 > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:b2_k_halo -s 4 -c 2 -o gpurun_out/halo_kernel python scripts/swe_steps.py 4096 3 >> gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/pytest_gpu_n1.log gpurun_out/smoke.log gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
