"""Turn ncu exports into the small JSON summaries committed under profiles/.

  python scripts/summarize_ncu.py raw  <raw.csv>  <out.json> "<capture command>" "<note>"
      raw.csv  = ncu -i X.ncu-rep --page raw --csv          (one row per captured launch)
  python scripts/summarize_ncu.py list <launches.csv> <out.json> "<command>" "<note>"
      launches.csv = ncu --metrics gpu__time_duration.sum --csv --log-file ...   (launch list)
"""
import csv
import json
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max",
]


def raw(path, out, command, note):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
    launches = []
    for r in rows[2:]:
        m = {k: {"value": r[idx[k]], "unit": units[idx[k]]} for k in METRICS if k in idx}
        st = sorted(((float(r[idx[k]]), k) for k in stall), reverse=True)[:8]
        m["stall_reasons_per_issue"] = {
            k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(v, 3)
            for v, k in st}
        launches.append({"kernel": r[idx["Kernel Name"]], "metrics": m})
    json.dump({"capture": command, "note": note, "launches": launches}, open(out, "w"), indent=1)


def launch_list(path, out, command, note):
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        name = r[i_name].split("(")[0].split("<")[0].split("::")[-1].strip()
        v = float(r[i_val].replace(",", ""))
        unit = r[i_unit]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    ks = [{"kernel": k, "launches": a[0], "total_ms": round(a[1], 3), "share": round(a[1] / tot, 5)}
          for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    json.dump({"command": command, "note": note, "total_ms": round(tot, 3), "kernels": ks}, open(out, "w"), indent=1)


if __name__ == "__main__":
    {"raw": raw, "list": launch_list}[sys.argv[1]](*sys.argv[2:6])
