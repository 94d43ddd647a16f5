#!/bin/bash
# multi-GPU evidence: byte-identical output for 1/2/4/8 devices + the strong-scaling bench line per N
# usage: tools/r2_multigpu.sh <max gpus> [skip-tests]   (writes gpurun_out/r2_multigpu_*)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_multigpu_gpus.txt 2>&1
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests/test_gpu.py -q -m gpu -k "multi_gpu" -rs > gpurun_out/r2_multigpu_pytest.txt 2>&1
  echo "pytest rc=$?" >> gpurun_out/r2_multigpu_pytest.txt
fi
for n in ${NLIST:-2 4 8}; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_multigpu_bench_n$n.json 2> gpurun_out/r2_multigpu_bench_n$n.err
    echo "bench n=$n rc=$?" >> gpurun_out/r2_multigpu_pytest.txt
  fi
done
