#!/bin/bash
# round 2, GPU pass C: tests, pool (dynamic claims), small-shard timing policy, probe experiments, ncu of the
# threshold / select kernels, ingest driver throughput
set -o pipefail
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -30 > gpurun_out/r02d_pytest.log; echo pytest $?
timeout 200 python tools/time_pool.py --json gpurun_out/r02d_time_pool.json > gpurun_out/r02d_time_pool.log 2>&1; echo pool $?
B="python bench.py --workloads none --no-cpu-baseline --sub-batches '' --steps 100 --warmup 5"
eval timeout 200 $B --workload c2s8 > gpurun_out/r02d_c2s8.json 2> gpurun_out/r02d_c2s8.err; echo c2s8 $?
ARCHI_TC_WARM=0 eval timeout 200 $B --workload c2s8 > gpurun_out/r02d_c2s8_warm0.json 2>/dev/null; echo c2s8w0 $?
eval timeout 200 $B --batch 64 > gpurun_out/r02d_c2_b64.json 2>/dev/null; echo b64 $?
ARCHI_TC_WARM=0 eval timeout 200 $B --batch 64 > gpurun_out/r02d_c2_b64_warm0.json 2>/dev/null; echo b64w0 $?
eval timeout 200 $B > gpurun_out/r02d_c2.json 2>/dev/null; echo c2 $?
ARCHI_TC_PROBE=24 eval timeout 200 $B > gpurun_out/r02d_c2_probe24.json 2>/dev/null; echo c2p24 $?
ARCHI_TC_PROBE=8 eval timeout 200 $B > gpurun_out/r02d_c2_probe8.json 2>/dev/null; echo c2p8 $?
NB="--kernel-name-base demangled"
Q='--no-cpu-baseline --sub-batches "" --parity 0'
eval timeout 250 ncu --set full --clock-control none --import-source on $NB -k '"regex:tc_maxima_threshold|tc_select"' -s 8 -c 2 -f -o gpurun_out/r02d_thr_select python bench.py --workloads none $Q --steps 1 --warmup 3 ">/dev/null" "2>gpurun_out/ncu_i.err"; echo ncuI $?
timeout 400 python tools/bench_ingest_driver.py > gpurun_out/r02d_ingest_driver.json 2> gpurun_out/r02d_ingest_driver.err; echo ingest $?
timeout 400 python tools/bench_ingest_driver.py --per-file > gpurun_out/r02d_ingest_driver_perfile.json 2>> gpurun_out/r02d_ingest_driver.err; echo ingestpf $?
tail -4 gpurun_out/r02d_pytest.log
cat gpurun_out/r02d_time_pool.log | cut -c1-420
cat gpurun_out/r02d_ingest_driver.json gpurun_out/r02d_ingest_driver_perfile.json
for f in c2s8 c2s8_warm0 c2_b64 c2_b64_warm0 c2 c2_probe24 c2_probe8; do python - <<P
import json
for l in open('gpurun_out/r02d_$f.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$f', round(d['ms_per_step'],4), round(d['value']), 'launch_ms', round(d['roofline']['launch_ms'],4), 'host', round(d.get('host_enqueue_ms_per_step',0),4), 'parity', d.get('parity_checked'), d.get('parity_failed'), 'unv', d.get('unverified_queries'), d['clocks']['reasons'])
P
done
