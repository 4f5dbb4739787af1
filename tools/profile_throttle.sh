#!/bin/bash
# soft throttle of long scans: parity with the throttle forced on, config-3 A/B, DRAM bytes of the main launch
ARCHI_TC_THROTTLE=1 timeout 300 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_tensor_modes.py -m gpu -q -x 2>&1 | tail -4
B="python bench.py --workload c3 --workloads none --no-cpu-baseline --sub-batches '' --steps 10 --warmup 3"
for t in auto 0; do
  if [ $t = auto ]; then eval timeout 150 $B > gpurun_out/r02t_c3_$t.json 2>/dev/null; else ARCHI_TC_THROTTLE=0 eval timeout 150 $B > gpurun_out/r02t_c3_$t.json 2>/dev/null; fi
  python - <<P
import json
for l in open('gpurun_out/r02t_c3_$t.json'):
    if l.startswith('{'):
        d=json.loads(l); print('c3 throttle=$t', round(d['ms_per_step'],3), round(d['value']), 'launch_ms', round(d['roofline']['launch_ms'],3), 'parity', d['parity_checked'], d['parity_failed'], d['clocks'])
P
done
timeout 200 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none --kernel-name-base demangled -k regex:tc_coarse_pair -s 2 -c 2 --csv --log-file gpurun_out/r02t_c3_dram.csv python bench.py --workload c3 --workloads none --no-cpu-baseline --sub-batches "" --parity 0 --steps 1 --warmup 3 > /dev/null 2>&1
grep -v "^==" gpurun_out/r02t_c3_dram.csv | cut -d, -f5,13- | tail -7
