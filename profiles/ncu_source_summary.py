"""Aggregate an `ncu --page source --csv` dump of one kernel: instruction and stall-sample
shares per SASS opcode and along the instruction stream (first launch in the file only)."""
import collections
import csv
import sys


def main(path, bins=24):
    rows = list(csv.reader(open(path)))
    heads = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
    hdr = rows[heads[0]]
    end = heads[1] - 1 if len(heads) > 1 else len(rows)
    data = [r for r in rows[heads[0] + 1:end] if len(r) == len(hdr)]
    si, ns, ie = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    tot_s = sum(int(r[ns]) for r in data)
    tot_i = sum(int(r[ie]) for r in data)
    print('SASS instructions', len(data), 'warp instructions executed', tot_i, 'samples', tot_s)
    byop, byop_s = collections.Counter(), collections.Counter()
    for r in data:
        parts = r[si].strip().split()
        op = parts[0] if not parts[0].startswith('@') else parts[1]
        op = op.split('.')[0]
        byop[op] += int(r[ie])
        byop_s[op] += int(r[ns])
    for op, c in byop.most_common(22):
        print('%-10s instr %5.1f%%  samples %5.1f%%' % (op, 100 * c / tot_i, 100 * byop_s[op] / tot_s))
    n = len(data)
    for b in range(bins):
        seg = data[b * n // bins:(b + 1) * n // bins]
        s = sum(int(r[ns]) for r in seg)
        i = sum(int(r[ie]) for r in seg)
        print('bin %2d sass %4d-%4d samples %5.1f%% instr %5.1f%%  first: %s' % (
            b, b * n // bins, (b + 1) * n // bins, 100 * s / tot_s, 100 * i / tot_i,
            seg[0][si].strip()[:60]))


if __name__ == '__main__':
    main(sys.argv[1])
