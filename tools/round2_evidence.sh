#!/bin/bash
# Round-2 evidence, one GPU box call:  gpurun --timeout 2400 -- 'bash tools/round2_evidence.sh'
# Bench lines are never taken under a profiler; ncu runs use --clock-control none.
set -u
O=gpurun_out/r02
mkdir -p $O
B="python bench.py --steps 5 --warmup 3"
$B > $O/bench_f64.json 2> $O/bench_f64.err
$B --impl reference > $O/bench_reference.json 2>> $O/bench_f64.err
$B --no-cpu-baseline --no-e2e --precision f32 > $O/bench_f32.json 2>> $O/bench_f64.err
QB_BP_MS2=0 $B --no-cpu-baseline --no-e2e > $O/bench_f64_round1_kernel.json 2>> $O/bench_f64.err
S="--no-cpu-baseline --e2e-shots 65536"
$B $S --schedule serial > $O/bench_serial_ms.json 2>> $O/bench_f64.err
$B $S --schedule serial --bp-method product_sum --osd-method osd_cs --osd-order 1 > $O/bench_serial_doc_setting.json 2>> $O/bench_f64.err
$B $S --bp-method product_sum > $O/bench_product_sum_flooding.json 2>> $O/bench_f64.err
$B $S --osd-method osd_cs --osd-order 1 > $O/bench_osd_cs1.json 2>> $O/bench_f64.err
$B $S --osd-method lsd_0 > $O/bench_lsd.json 2>> $O/bench_f64.err
$B $S --workload bb144_r10_p3e-3 > $O/bench_cfg3_p3e-3.json 2>> $O/bench_f64.err
$B $S --workload bb144_r10_p3e-4 > $O/bench_cfg3_p3e-4.json 2>> $O/bench_f64.err
python bench.py --steps 4 --warmup 3 $S --workload bb72_r6_p1e-3 > $O/bench_cfg2_bb72_1e6shots.json 2>> $O/bench_f64.err
python bench.py --steps 3 --warmup 3 $S --shots 65536 --workload hgp225_r3_p1e-2 --schedule serial --bp-method product_sum --osd-method osd_cs --osd-order 1 > $O/bench_cfg1_hgp225_doc_setting.json 2>> $O/bench_f64.err
python bench.py --steps 3 --warmup 3 $S --shots 65536 --workload hgp225_r15_p1e-3 --schedule serial --bp-method product_sum --osd-method osd_cs --osd-order 1 > $O/bench_hgp225_r15_doc06A_setting.json 2>> $O/bench_f64.err
python bench.py --steps 3 --warmup 3 $S --workload qt633_zxcol_r12_p1e-3 --osd-method lsd_0 > $O/bench_cfg4_qt633_bplsd.json 2>> $O/bench_f64.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --shots 16384 --workload qlp1020_zxcol_r20_p5e-4 > $O/bench_cfg5_qlp1020_osd0.json 2>> $O/bench_f64.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --shots 16384 --workload qlp1020_zxcol_r20_p5e-4 --schedule serial --bp-method product_sum --osd-method osd_cs --osd-order 1 > $O/bench_cfg5_qlp1020_doc_setting.json 2>> $O/bench_f64.err
# instruction model (warp instructions / shared wavefronts per edge-iteration of the BP launches of one batch)
M=smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,gpu__time_duration.sum
for p in f64 f32; do
  ncu --metrics $M --clock-control none -k regex:bp_kernel --csv --log-file $O/inst_$p.csv python bench.py --steps 1 --warmup 0 --shots 65536 --no-e2e --no-cpu-baseline --precision $p > $O/inst_${p}_bench.json 2> $O/inst_$p.err
done
# launch list of one step
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file $O/launches_f64.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2> $O/launches.err
# full captures: K3 (flooding min-sum), K1, serial slab kernel
ncu --set full --clock-control none --import-source on -k regex:bp_kernel_ms2 -s 4 -c 1 -o $O/bp_kernel_ms2_f64 python bench.py --steps 1 --warmup 1 --shots 65536 --no-e2e --no-cpu-baseline > /dev/null 2> $O/ncu1.err
ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 1 -c 1 -o $O/frame_kernel python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2> $O/ncu2.err
ncu --set full --clock-control none --import-source on -k regex:bp_kernel_serial_slab -s 4 -c 1 -o $O/bp_kernel_serial_slab_f64 python bench.py --steps 1 --warmup 1 --shots 65536 --no-e2e --no-cpu-baseline --schedule serial > /dev/null 2> $O/ncu3.err
ncu --set full --clock-control none --import-source on -k regex:bp_kernel_serial_slab -s 4 -c 1 -o $O/bp_kernel_serial_slab_ps_f64 python bench.py --steps 1 --warmup 1 --shots 65536 --no-e2e --no-cpu-baseline --schedule serial --bp-method product_sum > /dev/null 2> $O/ncu4.err
# sanitizer
QB_SANITIZE_SCALE=0.3 timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py > $O/compute_sanitizer_memcheck.log 2>&1
QB_SANITIZE_SCALE=0.1 timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py > $O/compute_sanitizer_racecheck.log 2>&1
tail -3 $O/compute_sanitizer_memcheck.log $O/compute_sanitizer_racecheck.log
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "%.4g" % d["value"], d.get("e2e", {}).get("value"), d.get("kernel_ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
