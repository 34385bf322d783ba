"""Turn ncu outputs under gpurun_out/ into the markdown summaries committed under profiles/.

    python tools/summarize_ncu.py full  gpurun_out/X.ncu-rep  profiles/X_ncu_full.md  "title / command"
    python tools/summarize_ncu.py list  gpurun_out/X_launches.csv profiles/X_launches.md "title / command"

`full`: one row per captured launch of an `ncu --set full --clock-control none` report (duration, tensor-pipe
activity, DRAM bytes read/written and % of peak, L2 bytes, warp instructions, registers, grid).
`list`: the `--metrics gpu__time_duration.sum` launch list aggregated per kernel name (launches, total, share).
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|st::|void ", "", name)
    name = re.sub(r"\((int|bool)\)", "", name)
    name = re.sub(r"\(CUtensorMap_st.*$|\(.*$", "", name) if len(name) > 90 else name
    return name[:90]


def col(hdr, key):
    for i, h in enumerate(hdr):
        if h == key:
            return i
    return None


def to_bytes(v, unit):
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v.replace(",", "")) * mul


def to_us(v, unit):
    mul = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)
    return float(v.replace(",", "")) * mul


def full(rep, out, title):
    if rep.endswith(".csv"):         # `ncu -i X.ncu-rep --page raw --csv` already run on the GPU box (reports over 64 MB do not travel)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = OrderedDict([
        ("time us", "gpu__time_duration.sum"),
        ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("DRAM rd MB", "dram__bytes_read.sum"),
        ("DRAM wr MB", "dram__bytes_write.sum"),
        ("DRAM %", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L2 sectors", "lts__t_sectors.sum"),
        ("warp insts", "smsp__inst_executed.sum"),
        ("regs", "launch__registers_per_thread"),
        ("smem KB", "launch__shared_mem_per_block_dynamic"),
    ])
    idx = {}
    for k, m in want.items():
        idx[k] = next((i for i, h in enumerate(hdr) if h == m), None)
    kn, gs, bs = col(hdr, "Kernel Name"), col(hdr, "Grid Size"), col(hdr, "Block Size")
    lines = [f"# {title}", "",
             f"Source: `{rep}` (`ncu --set full --clock-control none --import-source on`; per-launch numbers are "
             "cold-cache and serialised under the profiler -- the event-timed numbers of the same kernels are "
             "`roofline.kernels_ms` of the bench line next to this file).", "",
             "| # | kernel | grid | block | " + " | ".join(want) + " |",
             "|---:|---|---|---|" + "---:|" * len(want)]
    for n, r in enumerate(data):
        cells = []
        for k in want:
            i = idx[k]
            if i is None or i >= len(r) or r[i] == "":
                cells.append("-")
                continue
            u = units[i]
            if k == "time us":
                cells.append(f"{to_us(r[i], u):.1f}")
            elif k.endswith("MB"):
                cells.append(f"{to_bytes(r[i], u) / 1e6:.1f}")
            elif k == "smem KB":
                cells.append(f"{to_bytes(r[i], u) / 1e3:.1f}")
            elif k in ("warp insts", "L2 sectors"):
                cells.append(f"{int(float(r[i].replace(',', ''))):,}")
            else:
                try:
                    cells.append(f"{float(r[i].replace(',', '')):.1f}")
                except ValueError:
                    cells.append(r[i])
        lines.append(f"| {n} | `{short(r[kn])}` | {r[gs]} | {r[bs]} | " + " | ".join(cells) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def launch_list(path, out, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[0] == "ID":
            continue
        a = agg.setdefault(short(r[kn]), [0, 0.0])
        a[0] += 1
        a[1] += to_us(r[mv], r[mu])
    tot = sum(a[1] for a in agg.values())
    lines = [f"# {title}", "",
             f"Source: `{path}` (`ncu --metrics gpu__time_duration.sum --clock-control none`). All launches of the "
             "command (warm-up, CPU-side setup fills, L2 flushes and timed steps); times are serialised and "
             "cold-cache, so the SHARE column is what to compare with the event-timed `roofline.kernels_ms`.", "",
             f"Total: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time.", "",
             "| kernel | launches | total us | mean us | share % |", "|---|---:|---:|---:|---:|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {t:.1f} | {t / n:.2f} | {100 * t / tot:.1f} |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    (full if mode == "full" else launch_list)(src, dst, title)
