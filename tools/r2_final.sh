#!/bin/bash
# final single-GPU verification of the round: smoke, full GPU suite, default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.txt 2>&1
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_final_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_final_pytest.txt
timeout 900 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r2_final_bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2>> gpurun_out/r2_final_bench_n1.err
tail -2 gpurun_out/r2_final_smoke.txt; tail -3 gpurun_out/r2_final_pytest.txt; tail -2 gpurun_out/r2_final_bench_n1.err; cut -c1-400 gpurun_out/r2_final_bench_ref.json
