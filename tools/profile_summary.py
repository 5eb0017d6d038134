#!/usr/bin/env python
"""Turns gpurun_out/{launches_rNN.csv, prof_*_rNN.ncu-rep} into the tracked summaries under profiles/."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ncu_summary import WANT  # noqa: E402


def launch_table(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        n = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        tot[n] += v
        cnt[n] += 1
    T = sum(tot.values())
    out = ["| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        out.append("| `%s` | %d | %.3f | %.1f %% |" % (k, cnt[k], v, v / T * 100))
    return "\n".join(out), T


def raw_metrics(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    names = [r[hdr.index("Kernel Name")].split("(")[0] for r in rows[2:]]
    out.append("launches captured: " + ", ".join(names))
    out.append("")
    out.append("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(names))) + " |")
    out.append("|---|---|" + "---|" * len(names))
    vals = {}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out.append("| %s | %s | %s |" % (w, units[i], " | ".join(r[i] for r in rows[2:])))
            vals[w] = [r[i] for r in rows[2:]]
    return "\n".join(out), vals, units, hdr


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    src = os.path.join(ROOT, "gpurun_out")
    dst = os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    md = ["# ncu summaries, round %s (B200, `--clock-control none`)" % rnd[1:], "",
          "Command: `python bench.py --steps 2 --warmup 3 --pps 1 --no-e2e --no-cpu` (1920x1080, one pass per step).", ""]
    lp = os.path.join(src, "launches_%s.csv" % rnd)
    if os.path.exists(lp):
        table, T = launch_table(lp)
        open(os.path.join(dst, "launches_%s.csv" % rnd), "w").write(open(lp).read())
        md += ["## Launch list (`--metrics gpu__time_duration.sum`; cold-cache, serialised: compare SHARES)", "", table, "",
               "total %.3f ms over the captured launches" % T, ""]
    traffic = {}
    titles = {"trace3": "trace3 (k_trace on the 75 k-triangle scene of BASELINE config 3)", "confirmtma": "confirmtma (k_confirm_tma, the TMA A/B)",
              "neer": "neer (k_nee_resolve)"}
    for name, key in (("trace", "trace"), ("trace3", None), ("confirm", "confirm"), ("confirmtma", None), ("isaac", "isaac_raygen"),
                      ("shade", "shade_nee"), ("neer", None)):
        rep = os.path.join(src, "prof_%s_%s.ncu-rep" % (name, rnd))
        if not os.path.exists(rep):
            continue
        table, vals, units, hdr = raw_metrics(rep)
        md += ["## `--set full` capture: %s" % titles.get(name, name), "", table, ""]
        if key is None:
            continue
        try:
            def to_bytes(v, u):
                return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            traffic[key] = [to_bytes(a, units[ir]) + to_bytes(b, units[iw]) for a, b in zip(vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"])]
        except Exception as e:  # noqa
            pass
    open(os.path.join(dst, "ncu_%s.md" % rnd), "w").write("\n".join(md) + "\n")
    # per-launch DRAM traffic of the first captured launch of each kernel, for bench.py's roofline.traffic
    # (shade: the second capture is the NEE kernel, k_shade_surf<1>)
    json.dump({k: (v[1] if k == "shade_nee" and len(v) > 1 else v[0]) for k, v in traffic.items()}, open(os.path.join(dst, "traffic.json"), "w"), indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main()
