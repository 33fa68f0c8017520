#!/bin/bash
# round 2, call R (1 GPU): last check of the final tree -- examples / GEMM / model tests, smoke, bench defaults
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=30
timeout 200 python -m pytest tests/test_examples.py tests/test_gemm.py tests/test_models.py -q -m gpu -p no:cacheprovider > gpurun_out/r2r_pytest.log 2>&1
echo "pytest exit $?"; tail -n 2 gpurun_out/r2r_pytest.log | cut -c1-200
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2 | cut -c1-200
timeout 200 python bench.py > gpurun_out/r2r_bench_default.json 2> gpurun_out/r2r_bench_default.err
echo "bench exit $?"; python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r2r_bench_default.json') if l.startswith('{')][-1]
print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['checks'].get('checks_ok'))"
