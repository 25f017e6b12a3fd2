"""Turn the CSV exports of scripts/profile_capture.sh (gpurun_out/TAG_*.csv) into the committed summaries under
profiles/ (launch shares, per-layer tensor / memory metrics, DSP kernels) and write profiles/r02_traffic.json, which
bench.py reads for roofline.traffic.  Usage: python scripts/profile_summarise.py TAG OUTPREFIX "note"."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, outp, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def rows_of(path):
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    return list(csv.reader(lines))


def short(name):
    n = name.split("::")[-1]
    return n.split("(")[0].split("<")[0]


# ---- launch list (long format: one row per launch x metric) ----
lp = os.path.join(G, tag + "_launches.csv")
if os.path.exists(lp):
    r = rows_of(lp)
    hdr = r[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = {}
    for row in r[1:]:
        v = float(row[vi].replace(",", ""))
        ms = v / 1e6 if row[ui] in ("ns", "nsecond") else v / 1e3 if row[ui] in ("us", "usecond") else v
        a = tot.setdefault(short(row[ki]), [0, 0.0])
        a[0] += 1
        a[1] += ms
    allms = sum(a[1] for a in tot.values())
    with open(os.path.join(P, outp + "_launches_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv: python bench.py --steps 1 --warmup 1 --utts 32 --no-cpu-baseline (%s)\n" % note)
        for k, a in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write("%-40s launches=%4d total_ms=%9.3f share=%5.1f%%\n" % (k, a[0], a[1], 100 * a[1] / allms))
    with open(os.path.join(P, outp + "_launches.csv"), "w") as f:
        f.write(open(lp).read())

LAYERS = ["resblock1_1_conv2", "resblock1_2_conv1", "resblock1_2_conv2", "resblock2_1_conv1", "resblock2_1_conv2", "resblock2_2_conv1",
          "resblock2_2_conv2", "resblock3_1_conv1", "resblock3_1_conv2", "resblock3_2_conv1", "resblock3_2_conv2", "resblock4_1_conv1",
          "resblock4_1_conv2", "resblock4_2_conv1", "resblock4_2_conv2", "last_conv", "last_dense"]
COLS = ["launch__grid_size", "launch__cluster_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread"]


def wide(path):
    r = rows_of(path)
    return r[0], r[1], r[2:]


def cols(hdr, name):
    """All columns holding metric `name` (ncu repeats a metric once per section; only some are filled)."""
    exact = [i for i, h in enumerate(hdr) if h == name]
    suffix = [i for i, h in enumerate(hdr) if h != name and h.endswith("." + name)]
    return exact + suffix


def pick(row, idxs):
    for i in idxs:
        if row[i] != "":
            return i
    return idxs[0] if idxs else -1


def to_unit(v, u, want):
    v = float(v.replace(",", "")) if v else float("nan")
    scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
    if want == "ms" and u in scale:
        return v * scale[u]
    b = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}
    if want == "GB" and u in b:
        return v * b[u]
    return v


gp = os.path.join(G, tag + "_tc_raw.csv")
if os.path.exists(gp):
    hdr, units, rows = wide(gp)
    idx = [cols(hdr, c) for c in COLS]
    ki = hdr.index("Kernel Name")
    cmd = ("ncu --set full --clock-control none -k regex:gemm_shift|conv64_walk -s 31 -c 17 python bench.py --steps 1 --warmup 1 "
           "--utts 32 --no-cpu-baseline (%s; one 2048-window pass)" % note)
    per_kernel = {}
    with open(os.path.join(P, outp + "_tc_ncu_full.txt"), "w") as f:
        f.write(cmd + "; exported with --page raw --csv on the GPU box\n")
        f.write("columns: kernel, " + ", ".join(COLS) + "  [time ms, dram GB]\n")
        for n, row in enumerate(rows):
            name = LAYERS[n] if n < len(LAYERS) else "launch %d" % n
            vals = []
            for c, ii in zip(COLS, idx):
                i = pick(row, ii)
                if i < 0:
                    vals.append(float("nan"))
                    continue
                want = "ms" if "time_duration" in c else "GB" if "bytes" in c else ""
                vals.append(to_unit(row[i], units[i], want))
            kern = short(row[ki])
            f.write("%-20s %-20s" % (name, kern) + "".join("%12.4f" % v for v in vals) + "\n")
            if name.startswith("resblock"):
                per_kernel.setdefault(kern, {})[name] = vals[3] + vals[4]
    out = {"command": cmd, "kernels": {}}
    for kern, per in per_kernel.items():
        out["kernels"][kern] = {"dram_bytes_per_launch": 1e9 * sum(per.values()) / len(per), "launches_averaged": len(per),
                                "per_layer_GB": per}
    json.dump(out, open(os.path.join(P, "r02_traffic.json"), "w"), indent=1)
    with open(os.path.join(P, outp + "_tc_ncu_raw.csv"), "w") as f:
        f.write(open(gp).read())

dp = os.path.join(G, tag + "_dsp_raw.csv")
if os.path.exists(dp):
    hdr, units, rows = wide(dp)
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__grid_size", "launch__block_size", "launch__registers_per_thread"]
    ki = hdr.index("Kernel Name")
    with open(os.path.join(P, outp + "_dsp_ncu.txt"), "w") as f:
        f.write("ncu --set full --clock-control none -k regex:stft_kernel|istft_kernel -c 8: bench.py --steps 1 --warmup 1 --utts 256 (256 x 4 s; %s)\n" % note)
        for row in rows:
            parts = [short(row[ki]) + ("<phasor>" if "<1>" in row[ki] or "<(bool)1>" in row[ki] else "")]
            for w in want:
                i = pick(row, cols(hdr, w))
                if i >= 0:
                    parts.append("%s=%s%s" % (w, row[i], units[i]))
            f.write("  ".join(parts) + "\n")
print("wrote summaries to", P)
