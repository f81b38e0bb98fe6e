"""DRAM bytes per launch of the implicit-GEMM kernels (every `igemm*` instantiation, the dual launch included) from
an ncu launch list with dram__bytes_read.sum / dram__bytes_write.sum -> the json bench.py reads for roofline.traffic.
usage: igemm_traffic.py launches.csv out.json "<how the csv was made>" """
import csv
import json
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, per = None, {}
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if "igemm" not in d["Kernel Name"] or not d["Metric Name"].startswith("dram__bytes"):
            continue
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(d["Metric Unit"], 1)
        per[int(d["ID"])] = per.get(int(d["ID"]), 0.0) + v * scale
    total = sum(per.values())
    out = {"dram_bytes_per_launch": total / max(1, len(per)), "launches": len(per), "total_dram_bytes": total,
           "source": sys.argv[3] if len(sys.argv) > 3 else sys.argv[1]}
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
