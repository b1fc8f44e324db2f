"""Turns the ncu artefacts of a gpurun round into the markdown tables under profiles/.
    python tools/summarize_ncu.py launches gpurun_out/launches_r1c.csv
    python tools/summarize_ncu.py full gpurun_out/prof_r1c.ncu-rep"""
import csv, io, re, subprocess, sys
from collections import OrderedDict

mode, path = sys.argv[1], sys.argv[2]
if mode == "launches":
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("plagnn::", "")
        name = name.split("<unnamed>::")[-1]
        v = float(r[vi].replace(",", ""))
        if r[ui] == "ns": v /= 1e3
        elif r[ui] == "ms": v *= 1e3
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:70]}` | {c} | {t:.1f} | {t / tot:.3f} |")
    print(f"\ntotal {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
else:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    cols = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__cluster_size", "cluster"),
            ("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
            ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
            ("sm__inst_executed_pipe_tensor.sum", "tensor insts"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
            ("launch__registers_per_thread", "regs"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
    idx = [(hdr.index(c) if c in hdr else None, n) for c, n in cols]
    print("| " + " | ".join(n for _, n in idx) + " |\n|" + "---|" * len(idx))
    seen = set()
    for r in rows[2:]:
        key = (r[idx[0][0]][:50], r[idx[1][0]])
        if key in seen: continue
        seen.add(key)
        cells = []
        for i, n in idx:
            v = r[i] if i is not None else "n/a"
            if n == "kernel": v = "`" + re.sub(r"\(.*", "", v).replace("void ", "")[:44] + "`"
            else:
                try: v = f"{float(v.replace(',', '')):.1f}" if "." in v else v
                except ValueError: pass
            cells.append(v)
        print("| " + " | ".join(cells) + " |")
