"""Condense `ncu --page raw --csv` dumps into one markdown table: python tools/ncu_table.py file.csv [...]"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "IPC"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs")]


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
             and "not_issued" not in h]
    print(f"\n### {path}\n")
    print("| kernel | grid | " + " | ".join(k for _, k in KEYS) + " | GB/s (dram R+W / time) | top stalls (warps per issue) |")
    print("|---|---|" + "---|" * (len(KEYS) + 2))
    seen = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[-44:]
        grid = r[idx["Grid Size"]] if "Grid Size" in idx else ""
        key = (name, grid)
        seen[key] = seen.get(key, 0) + 1
        if seen[key] > 1:
            continue
        vals = []
        for k, _ in KEYS:
            v = num(r[idx[k]]) if k in idx else None
            u = units[idx[k]] if k in idx else ""
            if v is None:
                vals.append("")
                continue
            if k.startswith("gpu__time"):
                v = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
            if k.startswith("dram__bytes"):
                v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)   # MB
            vals.append(f"{v:.1f}" if v >= 10 else f"{v:.2f}")
        try:
            t_us, rd, wr = float(vals[0]), float(vals[1] or 0), float(vals[2] or 0)
            gbs = f"{(rd + wr) / t_us * 1e3:.0f}"
        except Exception:
            gbs = ""
        st = sorted(((num(r[i]) or 0, hdr[i].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                     for i in stall), reverse=True)[:3]
        print(f"| `{name}` | {grid} | " + " | ".join(vals) + f" | {gbs} | " + ", ".join(f"{n} {v:.2f}" for v, n in st) + " |")
