"""Key per-kernel counters of an .ncu-rep (raw page): usage ncu_key.py report.ncu-rep
NCU_SKIP=<n>: leave out the first n launches of the report (warm-up / set-up launches of the same kernels);
NCU_COUNT=<n>: read n launches."""
import csv, os, subprocess, sys, io
skip = (["--launch-skip", os.environ["NCU_SKIP"]] if os.environ.get("NCU_SKIP") else []) + \
       (["--launch-count", os.environ["NCU_COUNT"]] if os.environ.get("NCU_COUNT") else [])
out = subprocess.run(["ncu", "-i", sys.argv[1]] + skip + ["--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__bytes.sum.per_second', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_active.avg', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']
stalls = [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h]
for r in rows[2:]:
    print('=====')
    for w in want:
        if w in hdr:
            print('  %-72s %-16s %s' % (w, units[hdr.index(w)], r[hdr.index(w)][:80]))
    st = []
    for h in stalls:
        try:
            st.append((float(r[hdr.index(h)]), h))
        except ValueError:
            pass
    for v, h in sorted(st, reverse=True)[:8]:
        print('  %-72s %-16s %.3f' % (h, 'inst', v))
