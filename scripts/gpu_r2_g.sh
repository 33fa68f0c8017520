#!/bin/bash
# round 2, call G (1 GPU): the one-pass bulk kernel -- GPU suite, schedules, timeline, bench, per-launch times
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=30
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2g_pytest.log 2>&1
echo "pytest exit $?"; tail -n 6 gpurun_out/r2g_pytest.log | cut -c1-300
timeout 300 python scripts/swe_pipelines_bench.py 4096x4096 1024x2048 2048x2048 > gpurun_out/r2g_pipelines.log 2>&1
grep nx= gpurun_out/r2g_pipelines.log || tail -n 20 gpurun_out/r2g_pipelines.log
timeout 200 python scripts/swe_timeline.py 4096 6 > gpurun_out/r2g_timeline_4096.log 2>&1; tail -n 8 gpurun_out/r2g_timeline_4096.log
timeout 200 python scripts/swe_timeline.py 1448 6 > gpurun_out/r2g_timeline_1448.log 2>&1; tail -n 8 gpurun_out/r2g_timeline_1448.log
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 20 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err
echo "bench rc=$?"; cut -c1-420 gpurun_out/r2g_bench_n1.json; tail -n 3 gpurun_out/r2g_bench_n1.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench_n1_k20.json 2>> gpurun_out/r2g_bench_n1.err
cut -c1-200 gpurun_out/r2g_bench_n1_k20.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'swe_ca_bulk_step' -s 2 -c 2 \
  -o gpurun_out/r2g_bulk_step_full python scripts/swe_steps.py 4096 4 > gpurun_out/r2g_ncu_full.log 2>&1
tail -n 2 gpurun_out/r2g_ncu_full.log
python __graft_entry__.py --smoke > gpurun_out/r2g_smoke.log 2>&1; tail -n 2 gpurun_out/r2g_smoke.log
