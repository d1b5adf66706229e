"""Digest the CSV pages of an ncu report (exported on the GPU box with `ncu -i rep --page raw --csv` and
`--page source --csv`; the .ncu-rep itself is too big to bring back):
    python profiles/ncu_csv_digest.py <raw.csv> <source.csv> [hyps_per_launch]"""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__texin_sm2tex_req_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__tex_writeback_active.avg.pct_of_peak_sustained_elapsed', 'l1tex__f_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active', 'l1tex__t_requests_pipe_tex_mem_texture.sum',
        'l1tex__t_sectors_pipe_tex_mem_texture.sum', 'l1tex__m_xbar2l1tex_read_sectors.avg.pct_of_peak_sustained_elapsed',
        'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_shared_st.sum',
        'smsp__sass_inst_executed_op_global_ld.sum', 'smsp__sass_inst_executed_op_texture.sum',
        'sm__sass_thread_inst_executed_op_ffma_pred_on.sum', 'sm__sass_thread_inst_executed_op_fmul_pred_on.sum',
        'sm__sass_thread_inst_executed_op_fadd_pred_on.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_fmaheavy.sum', 'sm__inst_executed_pipe_fmalite.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_tex.sum', 'sm__inst_executed_pipe_uniform.sum',
        'sm__inst_executed_pipe_cbu.sum', 'sm__inst_executed_pipe_adu.sum', 'sm__inst_executed_pipe_fp64.sum']


def main():
    raw, src = sys.argv[1], sys.argv[2]
    hyps = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0 ** 20
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("kernel:", vals[hdr.index("Kernel Name")][:120])
    for i, h in enumerate(hdr):
        if h in WANT or ('issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(vals[i] or 0) > 0.05):
            print(f"{h:75s} {units[i]:10s} {vals[i]}")
    rows = list(csv.reader(open(src)))
    k = 0
    while 'Instructions Executed' not in rows[k]:
        k += 1
    hdr, data = rows[k], rows[k + 1:]
    iE, iT, iSm = hdr.index('Instructions Executed'), hdr.index('Avg. Threads Executed'), hdr.index('# Samples')
    iS = hdr.index('Source')
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
    # opcode histogram (dynamic)
    ops = {}
    for r in data:
        op = r[iS].split()[0] if r[iS].split() else '?'
        if op.startswith('@'):
            op = r[iS].split()[1]
        op = op.split('.')[0]
        ops[op] = ops.get(op, 0) + int(r[iE])
    print("dynamic opcode mix (warp instr per hypothesis):")
    print("  " + "  ".join(f"{o}:{c / hyps:.1f}" for o, c in sorted(ops.items(), key=lambda x: -x[1])[:24]))


if __name__ == "__main__":
    main()
