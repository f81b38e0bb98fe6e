"""Aggregate an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list
per kernel: launches, time, share of the step, DRAM bytes, achieved GB/s.  usage: launch_shares.py file.csv [top]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    hdr, per = None, collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ns": 1, "us": 1e3, "ms": 1e6}.get(d["Metric Unit"], 1)
        name = d["Kernel Name"].split("(")[0].replace("void ", "")
        per.setdefault((int(d["ID"]), name[:48]), {})[d["Metric Name"]] = v * scale
    ks = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
    for (_, k), m in per.items():
        a = ks[k]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0)
        a[2] += m.get("dram__bytes_read.sum", 0)
        a[3] += m.get("dram__bytes_write.sum", 0)
        a[4] += m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) * m.get("gpu__time_duration.sum", 0)
    tot = sum(a[1] for a in ks.values())
    print("| kernel | launches | time (ms) | share | DRAM read (GB) | DRAM write (GB) | DRAM GB/s | DRAM bytes / launch (MB) | tensor pipe % (time-weighted) |")
    print("|---|---|---|---|---|---|---|---|---|")
    for k, a in sorted(ks.items(), key=lambda kv: -kv[1][1])[:top]:
        if a[1] <= 0:
            continue
        print(f"| {k} | {a[0]} | {a[1] / 1e6:.2f} | {a[1] / tot:.3f} | {a[2] / 1e9:.2f} | {a[3] / 1e9:.2f} | "
              f"{(a[2] + a[3]) / a[1]:.0f} | {(a[2] + a[3]) / a[0] / 1e6:.1f} | {a[4] / a[1]:.1f} |")
    print(f"\ntotal kernel time {tot / 1e6:.2f} ms over {sum(a[0] for a in ks.values())} launches")


if __name__ == "__main__":
    main()
