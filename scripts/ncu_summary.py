"""Print the key metrics of an .ncu-rep (read here, no GPU needed): python scripts/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for row in rows[2:]:
    print('-' * 100)
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f'{w:82s} {row[i]:>24s} {units[i]}')

# optional: python scripts/ncu_summary.py file.ncu-rep traffic.json — per-kernel DRAM traffic of the captured layer
# (launch order: qkv, scores, softmax+P.V, wo, gate_up, down), read by bench.py for roofline.traffic
if len(sys.argv) > 2:
    import json
    names = ["qkv", "attn_scores", "attn_softmax_pv", "wo", "gate_up", "down"]
    ir, iw, it = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('gpu__time_duration.sum')
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out_j = {}
    for name, row in zip(names, rows[2:]):
        out_j[name] = {"kernel": row[hdr.index('Kernel Name')], "dram_bytes_read": float(row[ir]) * scale[units[ir]],
                       "dram_bytes_write": float(row[iw]) * scale[units[iw]], "ncu_duration_us": float(row[it]),
                       "source": sys.argv[1]}
    json.dump(out_j, open(sys.argv[2], "w"), indent=1)
