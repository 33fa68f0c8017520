#!/bin/bash
# round 2, call C (1 GPU): bench N=1 with the CA schedule; per-launch device times; ncu --set full of the CA kernels
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=20
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 20 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
echo "bench rc=$?"; cat gpurun_out/r2c_bench_n1.json; tail -n 5 gpurun_out/r2c_bench_n1.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c_bench_n1_k20.json 2>> gpurun_out/r2c_bench_n1.err
cat gpurun_out/r2c_bench_n1_k20.json
for n in 4096 1448; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/r2c_launches_$n.csv python scripts/swe_steps.py $n 3 > gpurun_out/r2c_ncu_launches_$n.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'swe_ca|halo_ca' -s 12 -c 6 \
  -o gpurun_out/r2c_swe_ca_full python scripts/swe_steps.py 4096 4 > gpurun_out/r2c_ncu_full.log 2>&1
tail -n 3 gpurun_out/r2c_ncu_full.log
