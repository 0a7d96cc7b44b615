"""Summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.csv"""
import csv, subprocess, sys, io
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active',
        'sass__inst_executed_local_loads', 'sm__cycles_elapsed.max']
keep += [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h]
with open(out, 'w') as f:
    wr = csv.writer(f)
    wr.writerow(['metric', 'unit'] + ['launch%d' % i for i in range(len(data))])
    for k in keep:
        if k in hdr:
            i = hdr.index(k)
            wr.writerow([k, units[i]] + [r[i] for r in data])
            print("%-90s %-10s %s" % (k, units[i], "  ".join(r[i] for r in data)))
