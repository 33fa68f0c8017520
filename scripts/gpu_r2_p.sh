#!/bin/bash
# round 2, call P (1 GPU): banded tile order of the tcgen05 GEMM
mkdir -p gpurun_out
timeout 200 python scripts/gemm_raster_sweep.py > gpurun_out/r2p_gemm_raster_sweep.log 2>&1
echo "rc=$?"; grep "^M=" gpurun_out/r2p_gemm_raster_sweep.log | cut -c1-700 || tail -n 20 gpurun_out/r2p_gemm_raster_sweep.log
timeout 200 python -m pytest tests/test_gemm.py -q -m gpu -p no:cacheprovider 2>&1 | tail -n 2
