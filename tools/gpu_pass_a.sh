#!/bin/bash
# round 2, GPU pass A: full GPU test suite, pool kernel timings + ncu, config-5 sub-bench, default bench wall time
set -o pipefail
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -40 > gpurun_out/r02b_pytest.log; echo pytest $?
timeout 200 python tools/time_pool.py --json gpurun_out/r02b_time_pool.json > gpurun_out/r02b_time_pool.log 2>&1; echo pool $?
timeout 300 python bench.py --workloads c5 --no-cpu-baseline --sub-batches "" --parity 0 --steps 5 --warmup 3 > gpurun_out/r02b_bench_c5.json 2> gpurun_out/r02b_bench_c5.err; echo c5 $?
NB="--kernel-name-base demangled"
timeout 200 ncu --set full --clock-control none --import-source on $NB -k regex:pool_ring -s 3 -c 1 -f -o gpurun_out/r02b_pool_ring python tools/time_pool.py --ring 1 --cases 1 > /dev/null 2> gpurun_out/ncu_f.err; echo ncuF $?
timeout 200 ncu --set full --clock-control none --import-source on $NB -k regex:pool_normalize -s 3 -c 1 -f -o gpurun_out/r02b_pool_cta python tools/time_pool.py --ring 0 --cases 1 > /dev/null 2> gpurun_out/ncu_g.err; echo ncuG $?
timeout 300 ncu --set full --clock-control none $NB -k 'regex:hyb_|scan_topk_kernel<float, 1' -s 30 -c 12 -f -o gpurun_out/r02b_hybrid python bench.py --workloads c5 --no-cpu-baseline --sub-batches "" --parity 0 --steps 3 --warmup 3 > /dev/null 2> gpurun_out/ncu_h.err; echo ncuH $?
( time timeout 1200 python bench.py > gpurun_out/r02b_bench_default.json 2> gpurun_out/r02b_bench_default.err ) 2> gpurun_out/r02b_bench_default.time; echo bench $?
cat gpurun_out/r02b_bench_default.time
tail -5 gpurun_out/r02b_pytest.log
cat gpurun_out/r02b_time_pool.log
