#!/bin/bash
# round 2, call I (2 GPUs): find the send/recv failure of call H
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=20
timeout 200 python -m mpi4jax_b200.run -n 2 --timeout 180 --output-dir gpurun_out/r2i_p2p -m pytest tests/collective_ops/test_send_and_recv.py \
   -q -m gpu -p no:cacheprovider -x -v > /dev/null 2>&1
echo "exit $?"; tail -n 60 gpurun_out/r2i_p2p/rank0.log | cut -c1-220
echo ---- rank 1; tail -n 30 gpurun_out/r2i_p2p/rank1.log | cut -c1-220
timeout 200 python -m mpi4jax_b200.run -n 2 --timeout 180 --output-dir gpurun_out/r2i_p2p_b -m pytest tests/collective_ops/test_sendrecv.py tests/test_extensions.py tests/test_coresidency.py \
   -q -m gpu -p no:cacheprovider -x -v > /dev/null 2>&1
echo "exit $?"; tail -n 40 gpurun_out/r2i_p2p_b/rank0.log | cut -c1-220
