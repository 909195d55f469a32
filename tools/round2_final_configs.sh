#!/bin/bash
# BASELINE configs 1, 2, 4 (and the doc/06A HGP setting) with the final code:  gpurun --timeout 600 -- 'bash tools/round2_final_configs.sh'
set -u
O=gpurun_out/r02f
mkdir -p $O
S="--no-cpu-baseline --e2e-shots 65536"
python bench.py --steps 4 --warmup 3 $S --workload bb72_r6_p1e-3 > $O/bench_cfg2_bb72_1e6shots.json 2> $O/bench.err
python bench.py --steps 2 --warmup 3 $S --shots 65536 --workload hgp225_r3_p1e-2 --schedule serial --bp-method product_sum --osd-method osd_cs --osd-order 1 > $O/bench_cfg1_hgp225_doc_setting.json 2>> $O/bench.err
python bench.py --steps 2 --warmup 3 $S --shots 65536 --workload hgp225_r15_p1e-3 --schedule serial --bp-method product_sum --osd-method osd_cs --osd-order 1 > $O/bench_hgp225_r15_doc06A_setting.json 2>> $O/bench.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --shots 131072 --workload qt633_zxcol_r12_p1e-3 --osd-method lsd_0 > $O/bench_cfg4_qt633_bplsd.json 2>> $O/bench.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "%.4g" % d["value"], d.get("e2e", {}).get("value"), d.get("kernel_ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
