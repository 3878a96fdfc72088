"""Summarise an `ncu --page source --csv` export: executed warp instructions by opcode, and the hottest lines."""
import collections, csv, re, sys

def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
    hdr = rows[hi]
    ci, si, ss = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)')
    ops = collections.Counter(); tot = 0; samples = collections.Counter(); stot = 0
    lines = []
    for r in rows[hi + 1:]:
        if len(r) <= ci or not r[ci]:
            continue
        try:
            n = int(float(r[ci])); s = int(float(r[ss] or 0))
        except ValueError:
            continue
        m = re.match(r'\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)', r[si])
        op = m.group(2) if m else '?'
        ops[op] += n; tot += n; samples[op] += s; stot += s
        lines.append((n, s, r[0], r[si]))
    print(f"total warp instructions {tot}, stall samples {stot}")
    for op, n in ops.most_common(top):
        print(f"  {op:10s} {n:12d} {100*n/tot:5.1f}%   samples {100*samples[op]/max(stot,1):5.1f}%")
    return lines

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
