"""`ncu -i X.ncu-rep --page raw --csv` -> compact JSON of the metrics the roofline discussion uses (one entry per launch)."""
import csv, json, sys
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
rows = list(csv.reader([l for l in open(sys.argv[1]) if not l.startswith("==")]))
hdr, units, data = rows[0], rows[1], rows[2:]
out = []
for r in data:
    d = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = (r[i] + (" " + units[i] if units[i] else "")).strip()
    out.append(d)
json.dump({"command": sys.argv[2] if len(sys.argv) > 2 else "", "what": sys.argv[3] if len(sys.argv) > 3 else "", "launches": out}, sys.stdout, indent=1)
