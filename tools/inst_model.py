#!/usr/bin/env python
"""profiles/bp_inst_model.json from one ncu pass over a single-batch bench step.

    (on the GPU box)  ncu --metrics smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,gpu__time_duration.sum \
                          --clock-control none -k regex:bp_kernel --csv --log-file gpurun_out/inst.csv \
                          python bench.py --steps 1 --warmup 0 --shots 65536 --no-e2e --no-cpu-baseline > gpurun_out/inst_bench.json
    (here)            python tools/inst_model.py gpurun_out/inst.csv gpurun_out/inst_bench.json f64_minimum_sum_parallel

The bench line (taken under the profiler: its times are ignored) supplies the edge-iterations those launches executed
(decoder_stats.bp_edge_iters); the ncu CSV supplies the warp-instructions and shared-memory wavefronts of the same launches.
"""
import csv
import json
import os
import sys


def main():
    csv_path, bench_path, key = sys.argv[1:4]
    rows = []
    with open(csv_path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        rows.append(r)
    tot = {}
    launches = set()
    for r in rows:
        launches.add(r["ID"])
        tot[r["Metric Name"]] = tot.get(r["Metric Name"], 0.0) + float(r["Metric Value"].replace(",", ""))
    line = None
    for ln in open(bench_path):
        ln = ln.strip()
        if ln.startswith("{"):
            line = json.loads(ln)
    ei = line["decoder_stats"]["bp_edge_iters"]
    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "bp_inst_model.json")
    try:
        model = json.load(open(out_path))
    except Exception:
        model = {}
    model[key] = {"warp_inst_per_edge_iter": tot["smsp__inst_executed.sum"] / ei,
                  "smem_wavefronts_per_edge_iter": tot.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 0.0) / ei,
                  "launches": len(launches), "edge_iters": ei, "warp_inst": tot["smsp__inst_executed.sum"],
                  "source": "ncu smsp__inst_executed.sum over the %d BP launches of one %d-shot step (%s), %s" % (
                      len(launches), int(line["config"].get("shots_per_step_per_gpu", 0)), os.path.basename(csv_path), line["config"]["workload"])}
    json.dump(model, open(out_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(model[key], indent=1))


if __name__ == "__main__":
    main()
