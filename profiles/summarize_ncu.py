"""Turn an .ncu-rep (read with `ncu -i ... --page raw --csv`) and a launch-list csv into the
small tracked summaries under profiles/.   python profiles/summarize_ncu.py <rep> <launches.csv> <tag> <kernel-json-name>"""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_read.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}


def main():
    rep, launches, tag, out_json = sys.argv[1:5]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units, data = rows[0], rows[1], rows[2:]
    summary = {"source": rep, "launches_profiled": len(data), "kernel": data[0][head.index("Kernel Name")], "metrics": {}}
    for k in KEYS:
        if k in head:
            i = head.index(k)
            vals = [float(r[i].replace(",", "")) for r in data if r[i] not in ("", "n/a")]
            if vals:
                summary["metrics"][k] = {"unit": units[i], "mean": sum(vals) / len(vals)}
    m = summary["metrics"]
    def val(k):
        return m[k]["mean"] * UNIT.get(m[k]["unit"], 1.0)
    summary["dram_bytes_per_launch"] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    summary["kernel_ms_under_ncu"] = val("gpu__time_duration.sum")
    # launch list: share of each kernel in the profiled window
    tot = {}
    order = []
    with open(launches) as f:
        lr = list(csv.reader(l for l in f if l.startswith('"')))
    h = lr[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    for r in lr[1:]:
        name = r[ki].split("(")[0]
        t = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ui], 1e-6)
        if name not in tot:
            order.append(name)
            tot[name] = [0, 0.0]
        tot[name][0] += 1
        tot[name][1] += t
    total = sum(v[1] for v in tot.values())
    summary["launch_list"] = {n: {"launches": tot[n][0], "total_ms": round(tot[n][1], 4),
                                  "share": round(tot[n][1] / total, 4)} for n in order}
    json.dump(summary, open(out_json, "w"), indent=1)
    print(json.dumps(summary, indent=1)[:3000])


if __name__ == "__main__":
    main()
