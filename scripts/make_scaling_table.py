#!/usr/bin/env python3
"""profiles/r02_scaling.md from the raw lines of the scaling runs (scripts/run_final_n1.sh, scripts/run_scaling.sh)"""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
TAGS = {1: "r02p", 2: "r02q", 4: "r02q", 8: "r02r"}
NAMES = {1: {"weak": "bench_n1", "strong": "strong_n1", "loh1_h50": "loh1_h50_n1", "cxx_weak": "cxx_weak_n1", "cxx_strong": "cxx_strong_n1"}}


def line(n, key):
    name = NAMES.get(n, {}).get(key, "%s_n%d" % (key, n))
    f = os.path.join(P, "%s_%s.json" % (TAGS[n], name))
    if not os.path.exists(f):
        return None
    try:
        return json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception:
        return None


def row(title, key, weak):
    base = line(1, key)
    cells = []
    for n in (1, 2, 4, 8):
        d = line(n, key)
        if d is None or base is None:
            cells.append("—")
            continue
        eff = d["value"] / (base["value"] * n) if weak else d["value"] / (base["value"] * n)
        extra = ""
        if "station_ok" in d:
            extra = ", station %s (%.1e)" % ("ok" if d["station_ok"] else "BAD", d["station_rel_diff_to_golden"])
        cells.append("%.2f Gpts/s, %.2f ms/step%s (%.0f %%)" % (d["value"], d["ms_per_step"], extra, 100 * eff if n > 1 else 100))
    return "| %s | %s |" % (title, " | ".join(cells))


out = ["# Scaling on one 8xB200 box (round 2)", "",
       "Raw lines: `profiles/r02p_*` (N=1, `scripts/run_final_n1.sh`), `r02q_*` (N=2, 4) and `r02r_*` (N=8) (`scripts/run_scaling.sh N TAG` under",
       "`gpurun --gpus N`).  Device time (CUDA events), max over ranks; in brackets the efficiency = rate / (N x the N=1 rate of the same",
       "row).  Each N ran on whatever box the pool handed out (SM clocks 1905-1965 MHz under the power cap), so a few per cent between",
       "columns is box-to-box variation.  Halo exchange: copy-engine pushes over CUDA IPC (`config.exchange` of the lines).", "",
       "| | N=1 | N=2 | N=4 | N=8 |", "|---|---|---|---|---|",
       row("weak: 2048x2048x(128 N), `bench.py` (the default line)", "weak", True),
       row("weak, C++ driver `host/slab_driver.C`", "cxx_weak", True),
       row("strong: 2048x2048x256, `bench.py --config strong`", "strong", False),
       row("strong, C++ driver", "cxx_strong", False),
       row("LOH.1-h50 (601x601x341, 1073 steps), `bench.py --config loh1-h50`", "loh1_h50", False), ""]
# topography lines
out += ["gaussianHill-rev.in (config 4: Cartesian 128x128x1900 + curvilinear 128x128x106 points; `scripts/check_topo_multigpu.py`, every rank",
        "bit-identical to the single-GPU run; the curvilinear block lives on rank 0, `balanced` = rank 0 owns fewer Cartesian planes):", ""]
for n in (2, 4, 8):
    for bal in ("", "_bal"):
        f = os.path.join(P, "%s_topo%s_n%d.log" % (TAGS[n], bal, n))
        if os.path.exists(f):
            for l in open(f):
                if l.startswith("TIMING"):
                    out.append("* N=%d%s: %s" % (n, " balanced" if bal else "", l.split(": ", 2)[2].strip() if l.count(": ") >= 2 else l.strip()))
out += ["", "What limits strong scaling: the z-marching kernel pays a 4-plane prologue per launch, so the two face-row launches of a slab (2 rows,",
        "7 plane steps each) cost about 0.9 ms per face and pass whatever the slab's thickness, and rank 0 carries the SBP closure rows (3.0 ms",
        "per step); at 32 planes per GPU (2048x2048x256 over 8) these fixed costs are about half of the step.  LOH.1-h50 over 8 GPUs is 43 planes",
        "of 605x605 points per GPU (2.5 ms per step): launch latencies and the same fixed costs."]
open(os.path.join(P, "r02_scaling.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
