#!/usr/bin/env python
"""Turns the csv of
  ncu --metrics <list below> -k regex:dm_step_kernel -s <skip> -c <n> --csv --clock-control none python bench.py ...
into profiles/step_metrics_<char>.json, the file bench.py reads for the roofline block (issue-slot %, SM-active %, counted fp32 FLOP per
Update per environment, DRAM traffic per launch).  usage: python tools/ncu_metrics_json.py <csv> <char> <envs> <updates_per_launch> <source tag>"""
import csv
import json
import os
import sys

METRICS = ("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,"
           "sm__cycles_active.avg,sm__cycles_elapsed.avg,sm__inst_executed.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,"
           "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum")


def main():
    path, char, envs, upl, tag = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    name_i, unit_i, val_i, id_i = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
    per = {}
    for r in rows[1:]:
        v = float(r[val_i].replace(",", ""))
        u = r[unit_i]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3}.get(u, 1.0)
        per.setdefault(r[name_i], []).append(v * scale)
    avg = {k: sum(v) / len(v) for k, v in per.items()}
    n = len(next(iter(per.values())))
    flop = avg["smsp__sass_thread_inst_executed_op_fadd_pred_on.sum"] + avg["smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"] + 2.0 * avg["smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"]
    out = {
        "source": tag, "launches_averaged": n, "envs": envs, "updates_per_launch": upl,
        "kernel_ms_under_ncu": 1e3 * avg["gpu__time_duration.sum"],
        "dram_bytes_per_launch": avg["dram__bytes_read.sum"] + avg["dram__bytes_write.sum"],
        "issue_slot_pct_of_peak": avg["sm__inst_executed.sum.pct_of_peak_sustained_elapsed"],
        "issue_active_pct_while_active": avg.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "sm_active_pct": 100.0 * avg["sm__cycles_active.avg"] / avg["sm__cycles_elapsed.avg"],
        "warp_instructions_per_launch": avg.get("smsp__inst_executed.sum"),
        "fp32_flop_per_launch": flop, "fp32_flop_per_update_per_env": flop / (envs * upl),
        "flop_source": "ncu smsp__sass_thread_inst_executed_op_{fadd,fmul,ffma}_pred_on.sum (executed thread-level fp32 operations of dm_step_kernel, FMA = 2; includes the lanes of padding links)",
    }
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "step_metrics_%s.json" % char)
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
