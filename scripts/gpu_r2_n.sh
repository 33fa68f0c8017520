#!/bin/bash
# round 2, call N (1 GPU): what the driver runs at round end -- GPU suite at world size 1, smoke(), bench.py
# defaults -- plus a fresh ncu capture of the shallow-water kernels
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=30
T0=$(date +%s)
stamp() { echo "== $1 (+$(( $(date +%s) - T0 )) s)"; }
stamp "pytest -m gpu"
timeout 400 python -m pytest tests/ -q -m gpu -p no:cacheprovider -x > gpurun_out/r2n_pytest_gpu_n1.log 2>&1
echo "pytest exit $?"; tail -n 4 gpurun_out/r2n_pytest_gpu_n1.log | cut -c1-300
stamp "smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2n_smoke.log 2>&1
echo "smoke exit $?"; tail -n 3 gpurun_out/r2n_smoke.log | cut -c1-300
stamp "bench"
timeout 300 python bench.py > gpurun_out/r2n_bench_default.json 2> gpurun_out/r2n_bench_default.err
echo "bench exit $?"; cut -c1-600 gpurun_out/r2n_bench_default.json; tail -n 2 gpurun_out/r2n_bench_default.err | cut -c1-200
timeout 100 python bench.py --impl reference 2>&1 | tail -n 1 | cut -c1-300
stamp "ncu"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'swe_ca_|halo_ca' -s 10 -c 5 \
  -o gpurun_out/r2n_swe_ca_step_full python scripts/swe_steps.py 4096 4 > gpurun_out/r2n_ncu_full.log 2>&1
tail -n 2 gpurun_out/r2n_ncu_full.log | cut -c1-200
stamp "done"
