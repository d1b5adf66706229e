"""profiles/<workload>_ncu_summary.json from the CSV pages of one `ncu --set full` capture (exported on the GPU
box, see profiles/r2_call2.sh):  python profiles/summarize_ncu_csv.py <raw.csv> <source.csv> <hyps> <out.json> [note]"""
import csv
import json
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__tex_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "smsp__thread_inst_executed_per_inst_executed.ratio"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}


def main():
    raw, src, hyps, out = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
    note = sys.argv[5] if len(sys.argv) > 5 else ""
    rows = list(csv.reader(open(raw)))
    head, units, data = rows[0], rows[1], rows[2]
    m = {}
    for k in KEYS:
        if k in head:
            i = head.index(k)
            if data[i] not in ("", "n/a"):
                m[k] = {"unit": units[i], "value": float(data[i].replace(",", ""))}
    val = lambda k: m[k]["value"] * UNIT.get(m[k]["unit"], 1.0)
    rows = list(csv.reader(open(src)))
    k = 0
    while "Instructions Executed" not in rows[k]:
        k += 1
    iE = rows[k].index("Instructions Executed")
    tot = sum(int(r[iE]) for r in rows[k + 1:])
    s = {"kernel": data[head.index("Kernel Name")], "hypotheses_per_launch": hyps, "note": note, "metrics": m,
         "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
         "kernel_ms_under_ncu": val("gpu__time_duration.sum"),
         "warp_instructions_per_hypothesis": tot / hyps}
    json.dump(s, open(out, "w"), indent=1)
    print(out, s["kernel"][:60], "dram MB", s["dram_bytes_per_launch"] / 1e6, "ms", s["kernel_ms_under_ncu"], "instr/hyp",
          round(s["warp_instructions_per_hypothesis"], 1))


if __name__ == "__main__":
    main()
