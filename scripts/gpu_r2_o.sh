#!/bin/bash
# round 2, call O (2 GPUs): ncu on rank 0 WITH a real peer -- single-pass metric list (the full set fails with
# UnknownError on kernels that touch peer-mapped memory: multi-pass replay cannot save / restore it)
mkdir -p gpurun_out
export MPI4JAX_B200_TIMEOUT=60
M=gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29662 --no-python \
   scripts/rank0_ncu.sh x --metrics $M --cache-control none -k regex:'b2_k_|swe_ca_' -s 14 -c 30 --csv --log-file gpurun_out/r2o_rank0_n2_metrics.csv -- \
   scripts/prof_collectives.py > gpurun_out/r2o_ncu_n2.log 2>&1
echo "ncu rc=$?"; grep -c "b2_k_\|swe_ca_" gpurun_out/r2o_rank0_n2_metrics.csv; tail -n 3 gpurun_out/r2o_ncu_n2.log | cut -c1-200
head -c 1500 gpurun_out/r2o_rank0_n2_metrics.csv
echo "== bench n=2 k=20"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2o_bench_n2_k20.json 2> gpurun_out/r2o_bench_n2_k20.err
cut -c1-250 gpurun_out/r2o_bench_n2_k20.json; python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r2o_bench_n2_k20.json') if l.startswith('{')][-1]
print(d['clocks'], d['e2e']['value'], d['checks'].get('checks_ok'))"
