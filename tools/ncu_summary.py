"""Condense an `ncu --page raw --csv` export into one row per launch with the metrics DESIGN.md argues from.
   python tools/ncu_summary.py raw.csv out.csv [first_launch]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
COLS = [
    ("time_us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs", "launch__registers_per_thread"),
    ("tensor_pipe_active_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("inst_executed", "smsp__inst_executed.sum"),
    ("pipe_alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("pipe_fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("pipe_xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("lsu_data_pipe_pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    ("tc_smem_wavefronts_pct", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("dram_read_bytes", "dram__bytes_read.sum"),
    ("dram_write_bytes", "dram__bytes_write.sum"),
    ("dram_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def val(r, name):
    if name not in ix:
        return ""
    v = r[ix[name]].replace(",", "")
    try:
        f = float(v)
    except ValueError:
        return v
    return f * SCALE.get(units[ix[name]], 1)


with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch", "kernel"] + [c for c, _ in COLS])
    tot_t = tot_tp = 0.0
    for i, r in enumerate(rows[2:]):
        if i < first:
            continue
        name = r[ix["Kernel Name"]].replace("void ", "").replace("cdm::", "").split("(CUtensorMap")[0][:70]
        vals = [val(r, m) for _, m in COLS]
        w.writerow([i, name] + [round(v, 3) if isinstance(v, float) else v for v in vals])
        if isinstance(vals[0], float) and isinstance(vals[4], float) and vals[4] > 0:
            tot_t += vals[0]
            tot_tp += vals[0] * vals[4]
    if tot_t:
        w.writerow(["time-weighted tensor_pipe_active_pct", round(tot_tp / tot_t, 2), "over", round(tot_t, 1), "us"])
print(open(sys.argv[2]).read()[-400:])
