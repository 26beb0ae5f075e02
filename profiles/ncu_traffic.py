"""DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) per kernel family from
`ncu --page raw --csv` dumps, written as the JSON bench.py reads for its `traffic` fields.

    python profiles/ncu_traffic.py out.json raw1.csv [raw2.csv ...]
"""
import csv
import json
import sys

FAMILIES = {                      # key prefix -> substring of the kernel name
    'fused_rows': 'rows_kernel<',
    'fused_columns': 'columns_cluster_kernel<',
    'fused_rows_fwd': 'rows_fwd_kernel<',
    'fused_columns_fwd': 'columns_fwd_cluster_kernel<',
    'grid_stage': 'grid_stage_kernel<',
    'grid_tma': 'grid_tma_kernel<',
    'degrid': 'degrid_kernel<',
    'clean_persistent': 'clean_persistent_kernel<',
}
UNITS = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main(out, paths):
    totals = {k: [0.0, 0] for k in FAMILIES}
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        name = hdr.index('Kernel Name')
        rd, wr = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
        for r in rows[2:]:
            for key, needle in FAMILIES.items():
                if needle in r[name]:
                    nbytes = (float(r[rd].replace(',', '')) * UNITS[units[rd]]
                              + float(r[wr].replace(',', '')) * UNITS[units[wr]])
                    totals[key][0] += nbytes
                    totals[key][1] += 1
    result = {'source': 'ncu --set full captures of `python bench.py --steps 1 --warmup 3 --no-cpu` '
                        '(profiles/r02_ncu_commands.sh, profiles/r02_*_ncu_summary.txt): '
                        'dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over '
                        'the captured launches'}
    for key, (nbytes, count) in totals.items():
        if count:
            result[key + '_dram_bytes_per_launch'] = nbytes / count
            result[key + '_launches_captured'] = count
    json.dump(result, open(out, 'w'), indent=1)
    print(json.dumps(result, indent=1))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2:])
