"""Summarise an .ncu-rep captured with --set full: key metrics per kernel and the top stalled SASS lines.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25] [--csv profiles/out.csv]
"""
import csv, io, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second',
        'lts__t_sector_op_read_hit_rate.pct', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def page(rep, name, extra=()):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv', *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 25
    rows = page(rep, 'raw')
    hdr, units = rows[0], rows[1]
    out_rows = [['Kernel Name'] + KEYS, [''] + [units[hdr.index(k)] if k in hdr else '' for k in KEYS]]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        print('=====', name)
        vals = []
        for k in KEYS:
            v = r[hdr.index(k)] if k in hdr else ''
            vals.append(v)
            print('  %-90s %s %s' % (k, v, units[hdr.index(k)] if k in hdr else ''))
        for h, v in zip(hdr, r):
            if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                try:
                    if float(v) > 0.1:
                        print('  stall %-60s %s' % (h.split('issue_stalled_')[1].split('_per_issue')[0], v))
                except ValueError:
                    pass
        out_rows.append([name] + vals)
    if '--csv' in sys.argv:
        with open(sys.argv[sys.argv.index('--csv') + 1], 'w', newline='') as f:
            csv.writer(f).writerows(out_rows)
    if top <= 0:
        return
    src = page(rep, 'source')
    # one block per kernel: "Kernel Name",<name> then header then lines
    i = 0
    while i < len(src):
        if src[i] and src[i][0] == 'Kernel Name':
            name = src[i][1]; h = src[i + 1]; i += 2
            data = []
            while i < len(src) and not (src[i] and src[i][0] == 'Kernel Name'):
                if len(src[i]) == len(h):
                    data.append(src[i])
                i += 1
            iS, iSrc, iEx = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
            sc = [j for j, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
            tot = sum(int(d[iS]) for d in data)
            print('===== source:', name[:80], 'samples', tot, 'sass lines', len(data))
            order = sorted(range(len(data)), key=lambda j: -int(data[j][iS]))[:top]
            for j in sorted(order):
                d = data[j]
                st = sorted(((h[c][6:], int(d[c])) for c in sc if int(d[c]) > 0), key=lambda kv: -kv[1])[:3]
                print('  %5d %6.2f%% ex=%-9s %-70s %s' % (j, 100.0 * int(d[iS]) / max(1, tot), d[iEx], d[iSrc].strip()[:70], st))
        else:
            i += 1


if __name__ == '__main__':
    main()
