#!/usr/bin/env python
"""ncu `--page raw --csv` export (tools/gpu_ncu.sh) -> a markdown table: per launch duration, DRAM bytes, tensor-pipe and
issue utilisation, instruction count, registers, shared memory.   python tools/ncu_summary.py gpurun_out/r02_ncu_raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
col = {h: j for j, h in enumerate(hdr)}


def val(r, name, scale_to=None):
    j = col[name]
    try:
        v = float(r[j].replace(",", ""))
    except ValueError:
        return float("nan")
    u = units[j]
    if scale_to == "MB":
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    if scale_to == "us":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    return v


print("| # | kernel | grid | time us | DRAM rd MB | DRAM wr MB | DRAM GB/s | tensor pipe % (elapsed) | issue active % | "
      "warp instr (M) | regs | dyn smem KB | smem bank conflicts (M) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for i, r in enumerate(data):
    if len(r) < len(hdr):
        continue
    t = val(r, "gpu__time_duration.sum", "us")
    rd, wr = val(r, "dram__bytes_read.sum", "MB"), val(r, "dram__bytes_write.sum", "MB")
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("vm::", "")
    print(f"| {i} | {name} | {r[col['Grid Size']]} | {t:.1f} | {rd:.1f} | {wr:.1f} | {(rd + wr) / t * 1e3:.0f} | "
          f"{val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{val(r, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{val(r, 'smsp__inst_executed.sum') / 1e6:.2f} | {int(val(r, 'launch__registers_per_thread'))} | "
          f"{val(r, 'launch__shared_mem_per_block_dynamic', 'MB') * 1e3:.0f} | "
          f"{val(r, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum') / 1e6:.2f} |")
