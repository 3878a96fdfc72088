"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per mesh call and kernel."""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = []
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki]
        m = re.match(r"(?:void )?(?:ctc::)?(\w+)(<[^>]*>)?", name)
        seq.append(((m.group(1) + (m.group(2) or "")) if m else name, float(r[vi].replace(",", ""))))
    runs, cur = [], None
    for s in seq:
        if "reset_state" in s[0]:
            cur = []
            runs.append(cur)
        if cur is not None:
            cur.append(s)
    for i, run in enumerate(runs):
        agg = collections.OrderedDict()
        for n, v in run:
            a = agg.setdefault(n, [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(a[1] for a in agg.values())
        print(f"--- mesh call {i}: {tot / 1e3:.1f} us over {len(run)} launches")
        for n, (c, v) in agg.items():
            print(f"   {n:42s} x{c:4d} {v / 1e3:10.1f} us {100 * v / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
