#!/bin/bash
# what the driver runs at round end (pytest -m gpu, smoke) + sanitizer evidence + p2p slot experiment
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/final_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/final_smoke.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/collective_ops/test_allreduce.py tests/collective_ops/test_sendrecv.py tests/collective_ops/test_alltoall.py tests/test_examples.py -q -m gpu -p no:cacheprovider -k "not ops_dtypes and not advanced" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck.log
for SLOT in 4194304 16777216; do
MPI4JAX_B200_P2P_SLOT_BYTES=$SLOT python -m mpi4jax_b200.run -n 2 --timeout 200 scripts/p2p_perf.py > gpurun_out/p2p_slot_$SLOT.log 2>&1
done
tail -n 3 gpurun_out/final_pytest_gpu.log gpurun_out/final_smoke.log
tail -n 6 gpurun_out/sanitizer_memcheck.log
grep -h "sendrecv" gpurun_out/p2p_slot_*.log
