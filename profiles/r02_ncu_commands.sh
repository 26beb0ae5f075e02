# ncu captures of round 2 (run on the GPU box through gpurun; see profiles/README.md).
# Reports are converted to CSV on the box and deleted (gpurun_out is limited to 64 MiB).
set -x
CMD="python bench.py --steps 1 --warmup 3 --no-cpu"
$CMD > gpurun_out/r2_plain.log 2>&1 || exit 1
# launch list of the whole run (3 warm-up steps + 1 timed step + the e2e leg + the side rows);
# takes ~12 minutes: SKIP_LIST=1 leaves it out
[ -n "$SKIP_LIST" ] || ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
capture() {   # name, kernel regex, skip, count, [kernel regex of the source page]
    ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/$1 $CMD > gpurun_out/$1.log 2>&1
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
    if [ -n "$5" ]; then ncu -i gpurun_out/$1.ncu-rep --page source --csv --kernel-name regex:"$5" > gpurun_out/$1_src.csv 2>/dev/null; fi
    rm -f gpurun_out/$1.ncu-rep
}
capture r2_g2i "rows_kernel|columns_cluster_kernel" 320 16 rows_kernel
[ -n "$SKIP_I2G" ] || capture r2_i2g "rows_fwd|columns_fwd" 80 8 ""
capture r2_grid "grid_tma|grid_stage|degrid_kernel|clear_columns|column_occupancy" 40 14 ""
capture r2_clean "clean_persistent|abs_histogram" 4 6 ""
du -sh gpurun_out
