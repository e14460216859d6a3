"""Summarise ncu captures into the tracked profiles/ directory (run in the build container, no GPU needed).

    python tools/ncu_summary.py launches gpurun_out/X_launches.csv profiles/X_launches.md [last_n] [batch]
        per-kernel table (launch count, total / mean device time, share) from a
        `ncu --metrics gpu__time_duration.sum --csv` launch list; `last_n` keeps only the last n launches
        (one forward) so that warm-up passes do not count.
    python tools/ncu_summary.py full gpurun_out/X.ncu-rep profiles/X_full.md
        the roofline-relevant metrics of every launch in a `ncu --set full` report.
    python tools/ncu_summary.py fullagg gpurun_out/X.ncu-rep profiles/X_full.md
        the same report grouped per kernel (launch count, total time, DRAM traffic and GB/s, time-weighted pipe / DRAM /
        SM utilisation, registers): one table for a whole forward captured with `--set full`.
"""
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    name = name.replace("s2f::", "")
    return name[:70]


def launches(src, dst, last_n=None, batch=None):
    """Per-kernel table from an ncu launch list; with dram__bytes_read/write.sum in the list it also reports the DRAM
    traffic and (batch given) refreshes profiles/traffic.json, which bench.py reads for roofline.traffic."""
    with open(src, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    per = {}                                   # launch id -> [name, us, dram bytes]
    order = []
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in csv.DictReader(io.StringIO("".join(lines))):
        lid = r["ID"]
        if lid not in per:
            per[lid] = [short(r["Kernel Name"]), 0.0, 0.0, False]
            order.append(lid)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            per[lid][1] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        elif r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            per[lid][2] += v * mult.get(unit, 1)
            per[lid][3] = True
    rows = [per[i] for i in order]
    if last_n:
        rows = rows[-int(last_n):]
    has_dram = any(r[3] for r in rows)
    agg = {}
    for name, us, by, _ in rows:
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += us
        a[2] += by
    tot = sum(a[1] for a in agg.values())
    totb = sum(a[2] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write(f"{len(rows)} launches, {tot / 1e3:.3f} ms of device time (ncu: cold-cache, serialised -- compare shares, not absolutes)")
        if has_dram:
            f.write(f"; DRAM traffic {totb / 1e9:.2f} GB" + (f" = {totb / 1e6 / int(batch):.0f} MB per image" if batch else ""))
        f.write("\n\n| kernel | launches | total us | mean us | share |" + (" DRAM read+write | per launch |" if has_dram else "") + "\n")
        f.write("|---|---:|---:|---:|---:|" + ("---:|---:|" if has_dram else "") + "\n")
        for name, (cnt, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {cnt} | {us:.1f} | {us / cnt:.1f} | {100 * us / tot:.1f}% |" +
                    (f" {by / 1e6:.0f} MB | {by / 1e6 / cnt:.1f} MB |" if has_dram else "") + "\n")
    print(open(dst).read())
    if has_dram and batch:
        import json
        import os
        g = [(c, us, by) for n, (c, us, by) in agg.items() if n.startswith("gemm_i8_tc_kernel")]
        if g:
            cnt, us, by = (sum(x[i] for x in g) for i in range(3))
            tj = os.path.join(os.path.dirname(os.path.abspath(dst)), "traffic.json")
            json.dump({"gemm_i8_tc_kernel<3>": {"launches_per_forward": cnt, "dram_bytes_per_launch": by / cnt,
                                                "share_of_step_under_ncu": us / tot, "batch": int(batch),
                                                "source": os.path.relpath(dst, os.path.dirname(os.path.dirname(os.path.abspath(dst))))}},
                      open(tj, "w"), indent=1)
            print("wrote", tj)


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for row in body:
            f.write(f"## `{short(row[idx['Kernel Name']])}`  grid {row[idx['Grid Size']]} block {row[idx['Block Size']]}\n\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"- {k} = {row[idx[k]]} {units[idx[k]]}\n")
            try:
                rd_b = float(row[idx["dram__bytes_read.sum"]].replace(",", ""))
                wr_b = float(row[idx["dram__bytes_write.sum"]].replace(",", ""))
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                tb = rd_b * mult.get(units[idx["dram__bytes_read.sum"]], 1) + wr_b * mult.get(units[idx["dram__bytes_write.sum"]], 1)
                t = float(row[idx["gpu__time_duration.sum"]].replace(",", ""))
                tu = {"nsecond": 1e-9, "ns": 1e-9, "usecond": 1e-6, "us": 1e-6, "msecond": 1e-3, "ms": 1e-3}.get(units[idx["gpu__time_duration.sum"]], 1e-9)
                f.write(f"- **traffic** (dram read+write) = {tb / 1e6:.2f} MB in {t * tu * 1e6:.1f} us -> {tb / (t * tu) / 1e9:.0f} GB/s\n")
            except Exception as e:  # noqa: BLE001
                f.write(f"- traffic: n/a ({e})\n")
            f.write("\n")
    print(open(dst).read())


def fullagg(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tmul = {"nsecond": 1e-9, "ns": 1e-9, "usecond": 1e-6, "us": 1e-6, "msecond": 1e-3, "ms": 1e-3}

    def num(row, k):
        try:
            return float(row[idx[k]].replace(",", ""))
        except Exception:  # noqa: BLE001
            return 0.0

    agg = {}
    for row in body:
        name = short(row[idx["Kernel Name"]])
        t = num(row, "gpu__time_duration.sum") * tmul.get(units[idx["gpu__time_duration.sum"]], 1e-9)
        by = num(row, "dram__bytes_read.sum") * mult.get(units[idx["dram__bytes_read.sum"]], 1) + \
            num(row, "dram__bytes_write.sum") * mult.get(units[idx["dram__bytes_write.sum"]], 1)
        a = agg.setdefault(name, dict(n=0, t=0.0, by=0.0, tensor=0.0, dram=0.0, sm=0.0, warps=0.0, regs=0, block=""))
        a["n"] += 1; a["t"] += t; a["by"] += by
        a["tensor"] += t * num(row, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        a["dram"] += t * num(row, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
        a["sm"] += t * num(row, "sm__throughput.avg.pct_of_peak_sustained_elapsed")
        a["warps"] += t * num(row, "sm__warps_active.avg.pct_of_peak_sustained_active")
        a["regs"] = int(num(row, "launch__registers_per_thread")); a["block"] = row[idx["Block Size"]]
    tot = sum(a["t"] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --set full, every launch of one forward, grouped per kernel ({src})\n\n")
        f.write(f"{sum(a['n'] for a in agg.values())} launches, {tot * 1e3:.3f} ms under ncu (cold caches, serialised: compare shares). "
                "Percentages are time-weighted means over the kernel's launches.\n\n")
        f.write("| kernel | launches | total us | share | DRAM MB | GB/s | tensor pipe active % | DRAM thr % | SM thr % | warps active % | regs | block |\n")
        f.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
            t = a["t"]
            f.write(f"| `{name}` | {a['n']} | {t * 1e6:.1f} | {100 * t / tot:.1f}% | {a['by'] / 1e6:.0f} | {a['by'] / t / 1e9:.0f} | "
                    f"{a['tensor'] / t:.1f} | {a['dram'] / t:.1f} | {a['sm'] / t:.1f} | {a['warps'] / t:.1f} | {a['regs']} | {a['block']} |\n")
    print(open(dst).read())


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None, sys.argv[5] if len(sys.argv) > 5 else None)
    elif mode == "fullagg":
        fullagg(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3])
