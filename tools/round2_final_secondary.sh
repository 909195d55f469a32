#!/bin/bash
# Secondary bench lines after the batch-capacity change:  gpurun --timeout 900 -- 'bash tools/round2_final_secondary.sh'
set -u
O=gpurun_out/r02d
mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B --schedule serial > $O/bench_serial_ms.json 2> $O/bench.err
$B --osd-method lsd_0 > $O/bench_lsd.json 2>> $O/bench.err
$B --osd-method lsd_cs --osd-order 1 > $O/bench_lsd_cs1.json 2>> $O/bench.err
$B --osd-method osd_cs --osd-order 1 > $O/bench_osd_cs1.json 2>> $O/bench.err
$B --bp-method product_sum --no-e2e > $O/bench_product_sum_flooding.json 2>> $O/bench.err
$B --rate 3e-3 --no-e2e > $O/bench_cfg3_p3e-3.json 2>> $O/bench.err
$B --rate 3e-4 --no-e2e > $O/bench_cfg3_p3e-4.json 2>> $O/bench.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "%.4g" % d["value"], d.get("e2e", {}).get("value"), d.get("kernel_ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
