#!/bin/bash
# round 2, call G2 (1 GPU): single-barrier bulk kernel + warp-per-quantity frame kernels
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=30
timeout 600 python -m pytest tests/test_examples.py tests/test_models.py tests/test_extensions.py tests/collective_ops/test_send_and_recv.py tests/collective_ops/test_sendrecv.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2g4_pytest.log 2>&1
echo "pytest exit $?"; tail -n 4 gpurun_out/r2g4_pytest.log | cut -c1-300
timeout 300 python scripts/swe_pipelines_bench.py 4096x4096 1024x2048 2048x2048 > gpurun_out/r2g4_pipelines.log 2>&1
grep nx= gpurun_out/r2g4_pipelines.log || tail -n 20 gpurun_out/r2g4_pipelines.log
timeout 200 python scripts/swe_timeline.py 4096 6 > gpurun_out/r2g4_timeline_4096.log 2>&1; tail -n 8 gpurun_out/r2g4_timeline_4096.log
timeout 200 python scripts/swe_timeline.py 1448 6 > gpurun_out/r2g4_timeline_1448.log 2>&1; tail -n 8 gpurun_out/r2g4_timeline_1448.log
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 20 > gpurun_out/r2g4_bench_n1.json 2> gpurun_out/r2g4_bench_n1.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/r2g4_bench_n1.json; tail -n 3 gpurun_out/r2g4_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'swe_ca_bulk_k12|swe_ca_bulk_fric|swe_ca_tend|swe_ca_fric_frame|halo_ca' -s 5 -c 10 \
  -o gpurun_out/r2g4_bulk_step_full python scripts/swe_steps.py 4096 4 > gpurun_out/r2g4_ncu_full.log 2>&1
tail -n 2 gpurun_out/r2g4_ncu_full.log
