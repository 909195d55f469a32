#!/bin/bash
# Round evidence on the GPU box (run through gpurun from the repo root): GPU tests, smoke, ncu launch list + full captures,
# bench lines.  Everything lands in gpurun_out/; the summaries are copied into profiles/ afterwards (tools/ncu_summary.py).
R=${1:-r01}
O=gpurun_out
B="python bench.py --steps 1 --warmup 1 --shots 65536 --no-e2e --no-cpu-baseline"
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/${R}_pytest_gpu.log; cat $O/${R}_pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches_f64.csv $B > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches_lsd.csv $B --osd-method lsd_0 --lanes 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bp_kernel_compact -s 5 -c 1 -o $O/${R}_bp_f64 -f $B > $O/ncu_bp.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lsd_kernel -s 1 -c 1 -o $O/${R}_lsd -f $B --osd-method lsd_0 --lanes 1 > $O/ncu_lsd.log 2>&1
if [ -n "$NCU_SERIAL" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:serial_warp -s 1 -c 1 -o $O/${R}_serial_warp -f python bench.py --steps 1 --warmup 1 --shots 8192 --no-e2e --no-cpu-baseline --schedule serial > $O/ncu_ser.log 2>&1
fi
timeout 400 python bench.py 2> $O/bench.err | tail -1 > $O/${R}_bench_f64.json
timeout 300 python bench.py --osd-method lsd_0 2>> $O/bench.err | tail -1 > $O/${R}_bench_lsd.json
timeout 300 python bench.py --schedule serial --shots 65536 --e2e-shots 65536 --steps 3 --warmup 3 2>> $O/bench.err | tail -1 > $O/${R}_bench_serial_ms.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>> $O/bench.err | tail -1 > $O/${R}_bench_reference.json
timeout 300 python bench.py --workload qlp1020_zxcol_r20_p5e-4 --shots 4096 --e2e-shots 4096 --steps 3 --warmup 3 2>> $O/bench.err | tail -1 > $O/${R}_bench_config5_qlp1020_osd0.json
timeout 300 python bench.py --workload qlp1020_zxcol_r20_p5e-4 --shots 4096 --e2e-shots 4096 --steps 3 --warmup 3 --osd-method lsd_0 --no-cpu-baseline 2>> $O/bench.err | tail -1 > $O/${R}_bench_config5_qlp1020_lsd0.json
tail -3 $O/bench.err
for f in f64 lsd serial_ms reference config5_qlp1020_osd0 config5_qlp1020_lsd0; do python -c "
import json,sys
d=json.load(open('$O/${R}_bench_$f.json'))
print('$f', round(d['value']), d.get('e2e',{}).get('value'), d.get('cpu_baseline',{}).get('value'), d.get('kernel_ms_per_step'))"; done
