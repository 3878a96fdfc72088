"""Turns the ncu exports of scripts/gpu_profiles_r2.sh (gpurun_out/) into the tracked summaries under profiles/."""
import collections, csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "scripts"))

KEYS = [("gpu__time_duration.sum", "time"), ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("smsp__inst_executed.sum", "warp inst"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %")]

def raw_table(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
        out.append((name, {lab: (r[idx[k]] + " " + units[idx[k]]).strip() for k, lab in KEYS if k in idx}))
    return out

def source_summary(path, top=14):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr = rows[hi]
    ci, si = hdr.index("Instructions Executed"), hdr.index("Source")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    ops, st, seen, tot = collections.Counter(), collections.Counter(), set(), 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[0] in seen:
            continue
        seen.add(r[0])
        try:
            n = int(float(r[ci]))
        except ValueError:
            continue
        m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[si])
        ops[m.group(2) if m else "?"] += n; tot += n
        for i in stall:
            try:
                st[hdr[i]] += int(float(r[i] or 0))
            except ValueError:
                pass
    s = sum(st.values()) or 1
    return tot, ops.most_common(top), [(k, 100.0 * v / s) for k, v in st.most_common(8)]

def main():
    lines = ["# ncu --set full, one mid-volume launch group (488 spans of the benched 1024^3 volume, fast mode, serial kernels) -- round 2",
             "", "`scripts/gpu_profiles_r2.sh` -> `gpurun_out/prof_r2_group.*` -> this file (`scripts/make_profiles_r2.py`). Per-launch values;",
             "the group is the fifth of nine (x = -0.15 .. 0.0 slab region: surface-heavy).", ""]
    tab = raw_table(os.path.join(G, "prof_r2_group.raw.csv"))
    labs = [lab for _, lab in KEYS]
    lines += ["| kernel | " + " | ".join(labs) + " |", "|---|" + "---|" * len(labs)]
    for name, d in tab:
        lines.append("| `" + name + "` | " + " | ".join(d.get(l, "") for l in labs) + " |")
    for title, f in (("K1 `sample_grids_kernel<fast, P8>`", "prof_r2_group.k1.source.csv"), ("E3 `vertex_kernel<fast, P8>`", "prof_r2_group.e3.source.csv")):
        tot, ops, st = source_summary(os.path.join(G, f))
        lines += ["", f"## {title}: executed warp instructions by opcode (source page), stall samples", "",
                  f"total {tot} warp instructions", "", "| opcode | count | share |", "|---|---|---|"]
        lines += [f"| {o} | {n} | {100.0 * n / tot:.1f} % |" for o, n in ops]
        lines += ["", "| stall reason | share of samples |", "|---|---|"] + [f"| {k} | {v:.1f} % |" for k, v in st]
    open(os.path.join(P, "ncu_group_full_r2.md"), "w").write("\n".join(lines) + "\n")
    # generic power
    tab = raw_table(os.path.join(G, "prof_r2_generic.raw.csv"))
    lines = ["# ncu --set full of `sample_grids_kernel<fast, generic>` (config 4: P = 4, 32 iterations, one mid-volume group) -- round 2", "",
             "The packed power-of-two path (`pow2_step`, two samples per thread, log2 P squarings unrolled).  The scalar trig-free step with",
             "run-time exponent loops it replaced for P = 2, 4, 16 took 7.41 ms for the same launch (issue 93 %, ALU pipe 66 %, FMA pipe 46 %);",
             "the cells of bench `other_configs.config4_power_sweep_1024cube` run 3.0-3.7x faster.", "",
             "| kernel | " + " | ".join(labs) + " |", "|---|" + "---|" * len(labs)]
    for name, d in tab:
        lines.append("| `" + name + "` | " + " | ".join(d.get(l, "") for l in labs) + " |")
    open(os.path.join(P, "ncu_generic_power_r2.md"), "w").write("\n".join(lines) + "\n")
    # K1 DRAM traffic per volume
    rows = list(csv.reader(l for l in open(os.path.join(G, "k1_traffic_r2.csv")) if l.startswith('"')))
    hdr = rows[0]; ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(float)
    for r in rows[1:]:
        k = "sample_grids" if "sample_grids" in r[ki] else "fixup_suspects"
        agg[(k, r[mi])] += float(r[vi].replace(",", ""))
    units = {r[mi]: r[hdr.index("Metric Unit")] for r in rows[1:]}
    def to_bytes(metric, v):
        u = units[metric].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    k1 = to_bytes("dram__bytes_read.sum", agg[("sample_grids", "dram__bytes_read.sum")]) + to_bytes("dram__bytes_write.sum", agg[("sample_grids", "dram__bytes_write.sum")])
    fx = to_bytes("dram__bytes_read.sum", agg[("fixup_suspects", "dram__bytes_read.sum")]) + to_bytes("dram__bytes_write.sum", agg[("fixup_suspects", "dram__bytes_write.sum")])
    json.dump({"dram_bytes_per_volume": k1 + fx, "sample_grids_kernel": k1, "fixup_suspects_kernel": fx,
               "algorithmic_bytes_per_volume": 4096 * (65 ** 3) * (4 + 1 / 8.0),
               "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the 9 + 9 launches of one volume (scripts/gpu_profiles_r2.sh)",
               "units_seen": units}, open(os.path.join(P, "k1_traffic_r2.json"), "w"), indent=1)
    print(open(os.path.join(P, "k1_traffic_r2.json")).read())

if __name__ == "__main__":
    main()
