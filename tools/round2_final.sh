#!/bin/bash
# Final evidence of round 2 after the batch-capacity change (262 144 shots per device batch, short first batch through the host
# arrays):  gpurun --timeout 1500 -- 'bash tools/round2_final.sh'
set -u
O=gpurun_out/r02c
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/pytest_gpu.log
B="python bench.py --steps 5 --warmup 3"
$B > $O/bench_f64.json 2> $O/bench.err
$B --impl reference > $O/bench_reference.json 2>> $O/bench.err
$B --no-cpu-baseline --precision f32 > $O/bench_f32.json 2>> $O/bench.err
$B --no-cpu-baseline --schedule serial --bp-method product_sum --osd-method osd_cs --osd-order 1 --steps 2 > $O/bench_serial_doc_setting.json 2>> $O/bench.err
M=smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,gpu__time_duration.sum
for p in f64 f32; do
  ncu --metrics $M --clock-control none -k regex:bp_kernel --csv --log-file $O/inst_$p.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --precision $p > $O/inst_${p}_bench.json 2> $O/inst_$p.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_f64.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2> $O/launches.err
cat $O/pytest_gpu.log
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "%.4g" % d["value"], d.get("e2e", {}).get("value"), d.get("kernel_ms_per_step"), d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
PY
