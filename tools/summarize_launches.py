"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total us, share).

    python tools/summarize_launches.py gpurun_out/launches.csv [skip_first_n | -last_n]

A negative second argument keeps only the LAST n launches: with `bench.py --quick --no-e2e --no-cpu-baseline` the list
ends with whole forwards (timed steps, then the event-profiled ones), so `-138` = two 69-launch forwards of the step.
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hdr, data = None, []
    for r in csv.reader(open(path)):
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        data.append(dict(zip(hdr, r)))
    data = data[skip:]      # (a negative value slices from the end)
    agg = collections.OrderedDict()
    for d in data:
        name = re.sub(r"\(.*", "", d["Kernel Name"])[:80]
        v = float(d["Metric Value"].replace(",", ""))
        v = v / 1000.0 if d["Metric Unit"] == "ns" else (v * 1000.0 if d["Metric Unit"] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{len(data)} launches, {tot:.1f} us total (cold-cache, serialised under ncu: compare SHARES, not absolutes)")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{a[0]:5d} x  {a[1]:10.1f} us  {a[1] / tot * 100:5.1f}%  avg {a[1] / a[0]:8.1f} us  {k}")


if __name__ == "__main__":
    main()
