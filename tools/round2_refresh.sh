#!/bin/bash
# Refresh of the headline evidence after the last kernel changes:  gpurun --timeout 1500 -- 'bash tools/round2_refresh.sh'
set -u
O=gpurun_out/r02b
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/pytest_gpu.log
B="python bench.py --steps 5 --warmup 3"
$B > $O/bench_f64.json 2> $O/bench.err
$B --impl reference > $O/bench_reference.json 2>> $O/bench.err
$B --no-cpu-baseline --no-e2e --precision f32 > $O/bench_f32.json 2>> $O/bench.err
QB_BP_MS2=0 $B --no-cpu-baseline --no-e2e > $O/bench_f64_round1_kernel.json 2>> $O/bench.err
$B --no-cpu-baseline --e2e-shots 65536 --osd-method lsd_0 > $O/bench_lsd.json 2>> $O/bench.err
M=smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,gpu__time_duration.sum
for p in f64 f32; do
  ncu --metrics $M --clock-control none -k regex:bp_kernel --csv --log-file $O/inst_$p.csv python bench.py --steps 1 --warmup 0 --shots 65536 --no-e2e --no-cpu-baseline --precision $p > $O/inst_${p}_bench.json 2> $O/inst_$p.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 40 --csv --log-file $O/launches_f64.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2> $O/launches.err
ncu --set full --clock-control none --import-source on -k regex:bp_kernel_ms2 -s 4 -c 1 -o $O/bp_kernel_ms2_f64 python bench.py --steps 1 --warmup 1 --shots 65536 --no-e2e --no-cpu-baseline > /dev/null 2> $O/ncu1.err
cat $O/pytest_gpu.log
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "%.4g" % d["value"], d.get("e2e", {}).get("value"), d.get("kernel_ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
