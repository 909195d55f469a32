#!/bin/bash
# Round-2 multi-GPU evidence:  gpurun --gpus 8 --timeout 1500 -- 'bash tools/round2_multi_gpu.sh 8'
set -u
N=${1:-8}
O=gpurun_out/r02
mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
nvidia-smi -L > $O/gpus_n$N.txt
# BASELINE config 3: 1e7 shots across the GPUs (5 steps x 262144 x N), all three noise rates, one process per GPU
$T 29601 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_f64_n$N.raw 2> $O/mg.err
$T 29602 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --workload bb144_r10_p3e-3 > $O/bench_cfg3_p3e-3_n$N.raw 2>> $O/mg.err
$T 29603 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --workload bb144_r10_p3e-4 > $O/bench_cfg3_p3e-4_n$N.raw 2>> $O/mg.err
# the reference's decoder on the same workload
$T 29604 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --e2e-shots 65536 --schedule serial --bp-method product_sum --osd-method osd_cs --osd-order 1 > $O/bench_serial_doc_setting_n$N.raw 2>> $O/mg.err
# BASELINE config 5: noise-rate sweep (circuits derived from the p = 5e-4 fixture by substituting the rate) with BP-OSD-0
for p in 3e-4 5e-4 1e-3; do
  $T 29605 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --shots 16384 --workload qlp1020_zxcol_r20_p5e-4 --rate $p > $O/bench_cfg5_qlp1020_p${p}_n$N.raw 2>> $O/mg.err
done
# one process, every GPU, through the drop-in calls
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --fanout --e2e-shots $((262144 * N)) > $O/bench_fanout_1proc_n$N.raw 2>> $O/mg.err
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fanout" 2>&1 | tail -2 > $O/pytest_fanout_n$N.log
for f in $O/*_n$N.raw; do grep "^{" $f | tail -1 > ${f%.raw}.json; done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/*_n$N.json")):
    try:
        d = json.load(open(f)); print(f.split("/")[-1], "%.4g" % d["value"], d.get("e2e", {}).get("value"), d.get("logical_errors"), d.get("shots_total"), d["clocks"])
    except Exception as e:
        print(f, "ERR", e)
PY
cat $O/pytest_fanout_n$N.log; tail -3 $O/mg.err
