"""Summarise .ncu-rep files into a markdown table (run where ncu is installed; no GPU needed).
usage: python tools/ncu_summary.py rep1.ncu-rep [rep2 ...] > profiles/xxx.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % peak"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    return hdr, units, r[2:]


def main():
    for rep in sys.argv[1:]:
        hdr, units, rows = rows_of(rep)
        idx = {h: i for i, h in enumerate(hdr)}
        print(f"\n### {rep.split('/')[-1]}\n")
        cols = [(m, n) for m, n in METRICS if m in idx]
        print("| kernel | " + " | ".join(f"{n} ({units[idx[m]]})" if units[idx[m]] else n for m, n in cols) + " | achieved GB/s |")
        print("|---|" + "---|" * (len(cols) + 1))
        for row in rows:
            name = row[idx["Kernel Name"]].split("(")[0]
            vals = [row[idx[m]] for m, _ in cols]

            def num(m):
                try:
                    return float(row[idx[m]].replace(",", ""))
                except Exception:
                    return float("nan")
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
            tscale = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}
            try:
                b = num("dram__bytes_read.sum") * scale[units[idx["dram__bytes_read.sum"]]] + \
                    num("dram__bytes_write.sum") * scale[units[idx["dram__bytes_write.sum"]]]
                t = num("gpu__time_duration.sum") * tscale[units[idx["gpu__time_duration.sum"]]]
                gbs = f"{b / t / 1e9:.0f}"
            except Exception:
                gbs = "?"
            print(f"| {name} | " + " | ".join(vals) + f" | {gbs} |")


if __name__ == "__main__":
    main()
