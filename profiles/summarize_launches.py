"""Per-kernel table from an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*]` CSV launch list.
    python profiles/summarize_launches.py gpurun_out/launches.csv [skip_fraction]
Times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import csv
import sys
from collections import OrderedDict


def main(path, skip=0.5):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H = rows[h]
    ki, mi, vi, ui, ii = (H.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    per = OrderedDict()
    for r in rows[h + 1:]:
        d = per.setdefault(int(r[ii]), {"name": r[ki]})
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        d[r[mi]] = v * scale
    ids = sorted(per)
    ids = ids[int(len(ids) * skip):]
    agg = OrderedDict()
    for i in ids:
        d = per[i]
        if "<unnamed>::" not in d["name"] or "at::" in d["name"]:
            key = "(torch plumbing kernels)"
        else:
            key = d["name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(key, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':44s} {'launches':>8s} {'us/launch':>10s} {'share':>7s} {'dram rd MB/l':>13s} {'dram wr MB/l':>13s} {'dram GB/s':>10s}")
    for k, (c, t, rd, wr) in agg.items():
        bw = (rd + wr) / (t * 1e-6) / 1e9 if t > 0 else 0.0
        print(f"{k[:44]:44s} {c:8d} {t / c:10.1f} {100 * t / tot:6.1f}% {rd / c / 1e6:13.2f} {wr / c / 1e6:13.2f} {bw:10.0f}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.5)
