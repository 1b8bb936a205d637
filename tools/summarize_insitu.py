"""Per-kernel DRAM traffic of a forward IN SITU (caches as the previous kernel left them), from an ncu CSV taken with

    ncu --cache-control none --clock-control none \
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file X.csv <bench command>

    python tools/summarize_insitu.py X.csv [-last_n_launches]

The default ncu captures flush the caches before every kernel, so they charge each kernel the HBM reads of operands
that, in the real step, the producing kernel has just left in the 126 MB L2.  This list answers how many bytes of the
16-bit activations' round trips between kernels actually reach HBM.
"""
import collections
import csv
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
        "nsecond": 1e-3, "msecond": 1e3}


def main():
    path = sys.argv[1]
    last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hdr, launches = None, collections.OrderedDict()
    for r in csv.reader(open(path)):
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        d = dict(zip(hdr, r))
        e = launches.setdefault(d["ID"], {"name": re.sub(r"\(.*", "", d["Kernel Name"])[:72]})
        e[d["Metric Name"]] = float(d["Metric Value"].replace(",", "")) * UNIT.get(d["Metric Unit"], 1.0)
    rows = list(launches.values())
    if last:
        rows = rows[last:]
    agg = collections.OrderedDict()
    for e in rows:
        a = agg.setdefault(e["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += e.get("gpu__time_duration.sum", 0.0)
        a[2] += e.get("dram__bytes_read.sum", 0.0)
        a[3] += e.get("dram__bytes_write.sum", 0.0)
    t = sum(a[1] for a in agg.values())
    rd = sum(a[2] for a in agg.values())
    wr = sum(a[3] for a in agg.values())
    print(f"{len(rows)} launches, {t:.1f} us, DRAM read {rd / 1e6:.1f} MB, write {wr / 1e6:.1f} MB "
          f"(in situ: --cache-control none; serialised under ncu)")
    print(f"{'n':>5}  {'avg us':>8}  {'read MB/launch':>14}  {'write MB/launch':>15}  {'GB/s':>7}  kernel")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        n = a[0]
        gbs = (a[2] + a[3]) / 1e9 / (a[1] * 1e-6) if a[1] > 0 else 0.0
        print(f"{n:5d}  {a[1] / n:8.1f}  {a[2] / n / 1e6:14.2f}  {a[3] / n / 1e6:15.2f}  {gbs:7.0f}  {k}")


if __name__ == "__main__":
    main()
