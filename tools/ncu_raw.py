"""Key metrics per launch from an .ncu-rep (ncu --page raw --csv)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'sm__inst_executed.sum',
        'smsp__cycles_active.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.avg.per_cycle_active']
ki = hdr.index('Kernel Name')
for r in rows[2:]:
    name = r[ki].split('(')[0].replace('void ', '')
    print('== ' + name)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print('   %-66s %14s %s' % (w, r[i], units[i]))
