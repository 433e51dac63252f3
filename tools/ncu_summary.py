#!/usr/bin/env python
"""Turn an ncu report (.ncu-rep, `--set full`) into a markdown table of the metrics the roofline uses.

    python tools/ncu_summary.py gpurun_out/ev/ncu_elem.ncu-rep [title] > profiles/rNN_xxx.md

One row per captured launch: duration, grid/block/regs, DRAM read/write bytes (the `traffic` figure),
DRAM throughput %, tensor-pipe %, L2->SM fill bytes, L2 hit rate, achieved occupancy.
Runs on the CPU box (ncu -i ... --page raw --csv).
"""
import csv
import io
import subprocess
import sys

COLS = [
    ("time_us", "gpu__time_duration.sum", 1.0),
    ("grid", "launch__grid_size", 1.0),
    ("block", "launch__block_size", 1.0),
    ("regs", "launch__registers_per_thread", 1.0),
    ("dram_rd_MB", "dram__bytes_read.sum", None),
    ("dram_wr_MB", "dram__bytes_write.sum", None),
    ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("tensor_%", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("sm_%", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("l2->sm_MB", "l1tex__m_xbar2l1tex_read_bytes.sum", None),
    ("l2_hit_%", "lts__t_sector_hit_rate.pct", 1.0),
    ("occ_%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
]
UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, data = rows[0], rows[1], rows[2:]

    def find(metric):
        for i, h in enumerate(head):
            if h == metric or h.endswith("." + metric):
                return i
        return None

    kname = find("Kernel Name")
    print(f"# {title}\n")
    print(f"source: `{rep}` (ncu --set full --clock-control none; per-launch, cold cache, serialised)\n")
    print("| kernel | " + " | ".join(c[0] for c in COLS) + " | traffic_MB |")
    print("|---|" + "---|" * (len(COLS) + 1))
    for r in data:
        cells, rd, wr = [], 0.0, 0.0
        for name, metric, scale in COLS:
            i = find(metric)
            if i is None or r[i] == "":
                cells.append("-")
                continue
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                cells.append("-")
                continue
            u = units[i]
            if scale is None:
                v *= UNIT.get(u, 1.0)
            elif name == "time_us":
                v *= TIME.get(u, 1.0)
            if name == "dram_rd_MB":
                rd = v
            if name == "dram_wr_MB":
                wr = v
            cells.append(f"{v:.1f}" if abs(v) < 1e5 and v != int(v) else f"{v:.0f}")
        k = r[kname].split("(")[0][-60:]
        print(f"| `{k}` | " + " | ".join(cells) + f" | {rd + wr:.1f} |")


if __name__ == "__main__":
    main()
