#!/usr/bin/env python
"""Summarise .ncu-rep files (read with `ncu -i ... --page raw --csv`) into a small markdown table for profiles/.

    python scripts/ncu_summary.py gpurun_out/a.ncu-rep [b.ncu-rep ...] > profiles/rN_ncu_summary.md
"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    u = unit.lower()
    return f * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


def to_ms(v, unit):
    f = float(v.replace(",", ""))
    return f * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "s": 1e3, "second": 1e3, "nsecond": 1e-6}.get(unit, 1)


def main():
    print("| report | # | kernel | time ms | DRAM rd MB | DRAM wr MB | DRAM GB/s | dram% | tensor% | hmma% | sm% | issue% | lsu% | warps% | regs | grid x block |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        for n, r in enumerate(rows[2:]):
            def g(k):
                i = ix.get(k)
                return (r[i], units[i]) if i is not None and r[i] != "" else (None, None)
            name = r[ix["Kernel Name"]]
            name = name.split("(")[0].replace("void ", "").replace("mpl::", "").replace("<unnamed>::", "")[:60]
            t, tu = g("gpu__time_duration.sum")
            ms = to_ms(t, tu) if t else float("nan")
            rd, ru = g("dram__bytes_read.sum")
            wr, wu = g("dram__bytes_write.sum")
            rdb = to_bytes(rd, ru) if rd else float("nan")
            wrb = to_bytes(wr, wu) if wr else float("nan")
            gbs = (rdb + wrb) / (ms * 1e-3) / 1e9
            vals = []
            for k, _ in COLS[3:11]:
                v, _u = g(k)
                vals.append(f"{float(v.replace(',', '')):.1f}" if v else "-")
            grid, _ = g("launch__grid_size")
            block, _ = g("launch__block_size")
            print(f"| {rep.split('/')[-1]} | {n} | `{name}` | {ms:.3f} | {rdb / 1e6:.1f} | {wrb / 1e6:.1f} | {gbs:.0f} | "
                  + " | ".join(vals) + f" | {grid} x {block} |")


if __name__ == "__main__":
    main()
