#!/bin/bash
# round 2, second GPU call (2 GPUs): partitioned variants (tests + full-size bench), chain-length sweep of the full-size gradient parity
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/r2_pytest_dist.log 2>&1
echo "pytest dist exit $?"; tail -15 gpurun_out/r2_pytest_dist.log
for c in 40 16 8; do
  PLAGNN_TMA_LONG_CHAIN=$c timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -p no:cacheprovider -k full_size > gpurun_out/r2_fullsize_chain$c.log 2>&1
  echo "full-size parity chain=$c exit $?"; cp gpurun_out/grad_parity_shared_full_size.txt gpurun_out/r2_grad_parity_full_chain$c.txt
done
timeout 600 python bench.py --workload scaled --steps 5 --warmup 3 > gpurun_out/r2_scaled_n1.json 2> gpurun_out/r2_scaled_n1.err; echo "scaled n1 exit $?"; cut -c1-600 gpurun_out/r2_scaled_n1.json; tail -5 gpurun_out/r2_scaled_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --workload scaled --steps 5 --warmup 3 > gpurun_out/r2_scaled_n2.json 2> gpurun_out/r2_scaled_n2.err; echo "scaled n2 exit $?"; cut -c1-600 gpurun_out/r2_scaled_n2.json; tail -5 gpurun_out/r2_scaled_n2.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-partitioned --no-cpu-baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench exit $?"; cut -c1-300 gpurun_out/r2_bench2.json
