"""Print the key metrics and the top stall reasons of every kernel in an ncu raw CSV
(`ncu -i X.ncu-rep --page raw --csv > X_raw.csv`)."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__grid_size',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    stall = [h for h in hdr if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]
    for r in rows[2:]:
        print(r[ki][:90])
        for w in WANT:
            if w in hdr:
                print('   {:75s} {} {}'.format(w, r[hdr.index(w)], units[hdr.index(w)]))
        vals = []
        for h in stall:
            v = r[hdr.index(h)].replace(',', '')
            vals.append((float(v) if v not in ('', 'n/a') else 0.0, h))
        total = sum(v for v, _ in vals) or 1.0
        for v, h in sorted(vals, reverse=True)[:8]:
            print('      stall {:5.1f} %  {}'.format(
                100 * v / total, h.replace('smsp__pcsamp_warps_issue_stalled_', '')))


if __name__ == '__main__':
    main(sys.argv[1])
