"""Digest an .ncu-rep into (a) headline raw metrics and (b) code regions by executed-instruction
count with their stall samples.   python profiles/ncu_digest.py <rep> [hyps_per_launch]"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    hyps = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0 ** 20
    rows = list(csv.reader(run([rep, "--page", "raw", "--csv"]).splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    for i, h in enumerate(hdr):
        if h in WANT or ('issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(vals[i] or 0) > 0.05):
            print(f"{h:75s} {units[i]:10s} {vals[i]}")
    rows = list(csv.reader(run([rep, "--page", "source", "--csv"]).splitlines()))
    hdr, data = rows[1], rows[2:]
    iE, iT, iSm = hdr.index('Instructions Executed'), hdr.index('Avg. Threads Executed'), hdr.index('# Samples')
    tot = sum(int(r[iE]) for r in data)
    tots = sum(int(r[iSm]) for r in data)
    print(f"total warp instructions {tot}  per hypothesis {tot / hyps:.1f}  samples {tots}")
    seg, cur = [], None
    for k, r in enumerate(data):
        e, s = int(r[iE]), int(r[iSm])
        t = float(r[iT]) if r[iT] not in ('', '-') else 0.0
        if cur and abs(cur['e'] - e) <= 0.02 * max(cur['e'], 1) + 50:
            cur['n'] += 1; cur['sum'] += e; cur['samples'] += s; cur['thr'] += t * e
        else:
            if cur:
                seg.append(cur)
            cur = {'start': k, 'e': e, 'n': 1, 'sum': e, 'samples': s, 'thr': t * e}
    seg.append(cur)
    print('  idx n_instr  exec_each  total(M)  share avg_thr samples sample_share')
    for s in seg:
        if s['sum'] > tot * 0.01 or s['samples'] > tots * 0.01:
            print(f"{s['start']:5d} {s['n']:6d} {s['e']:10d} {s['sum'] / 1e6:9.2f} {s['sum'] / tot:6.3f} "
                  f"{s['thr'] / max(s['sum'], 1):6.1f} {s['samples']:7d} {s['samples'] / max(tots, 1):6.3f}")


if __name__ == "__main__":
    main()
