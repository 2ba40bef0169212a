#!/usr/bin/env python3
"""Turn the raw ncu artefacts in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarise_ncu.py launches gpurun_out/launches_v11.csv profiles/r01_launches_bench_v11.csv
    python tools/summarise_ncu.py full gpurun_out/prof_plane_ws_v11.ncu-rep profiles/r01_ncu_full_k_plane_gain_ws_v11.txt \
           [pairs_in_launch [kernel-name regex, for a report that holds several kernels]]

`launches` copies the per-launch duration list (gpu__time_duration.sum, --clock-control none) and
prints/returns per-kernel shares; `full` extracts the metrics the roofline discussion uses from one
`ncu --set full` capture (read with `ncu -i ... --page raw/source --csv`).  Both also refresh the
matching entries of profiles/r02_ncu_summary.json ($BFSM_NCU_SUMMARY), which bench.py reads for
`roofline.traffic`.
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUMMARY = os.path.join(ROOT, "profiles", os.environ.get("BFSM_NCU_SUMMARY", "r02_ncu_summary.json"))
AIDS = ("k_dfma_peak", "at::", "vectorized_elementwise", "elementwise_kernel")  # measurement aids


def load_summary():
    try:
        with open(SUMMARY) as fh:
            return json.load(fh)
    except Exception:
        return {}


def save_summary(d):
    with open(SUMMARY, "w") as fh:
        json.dump(d, fh, indent=1)
        fh.write("\n")


def short_name(full):
    m = re.search(r"(?:bfsm::)?([A-Za-z_0-9:]+)\s*<", full) or re.search(r"([A-Za-z_0-9:]+)\(", full)
    name = m.group(1) if m else full
    return name.replace("bfsm::", "").replace("void ", "").lstrip(":")


def launches(src, dst, command):
    rows = [l for l in open(src) if l.startswith('"')]
    with open(dst, "w") as fh:
        fh.writelines(rows)
    rd = csv.DictReader(rows)
    per = {}
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = short_name(r["Kernel Name"])
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        per.setdefault(name, []).append(ns)
    total = sum(sum(v) for k, v in per.items() if not any(a in k for a in AIDS))
    kernels = {}
    for k, v in per.items():
        aid = any(a in k for a in AIDS)
        kernels[k] = {"launches": len(v), "mean_us": round(sum(v) / len(v) / 1e3, 2),
                      "share": None if aid else round(sum(v) / total, 4)}
    d = load_summary()
    d["command"] = command
    d["note"] = ("first 400 launches (cold-cache, serialised under ncu): compare SHARES, not absolutes; "
                 "shares exclude the measurement aids (k_dfma_peak = FP64 peak micro-benchmark, torch fill = L2 flush)")
    d["kernels"] = kernels
    save_summary(d)
    print(json.dumps(kernels, indent=1))


def ncu_csv(rep, page, kernel=None):
    cmd = ["ncu", "-i", rep, "--page", page, "--csv"]
    if kernel:
        cmd += ["-k", "regex:" + kernel]   # reports that hold several kernels: pick one
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def full(rep, dst, pairs, kernel_filter=None):
    raw = ncu_csv(rep, "raw", kernel_filter)
    hdr, vals = raw[0], raw[2]
    m = dict(zip(hdr, vals))
    kernel = short_name(m.get("Kernel Name", "?"))

    def f(key, default=0.0):
        try:
            return float(m[key].replace(",", ""))
        except Exception:
            return default

    unit = dict(zip(hdr, raw[1]))

    def scaled(key):
        v = f(key)
        u = unit.get(key, "")
        return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1.0)

    stalls = {}
    for h in hdr:
        mm = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active\.ratio", h)
        if mm:
            stalls[mm.group(1)] = f(h)
    tot = sum(stalls.values()) or 1.0
    stall_pct = {k: round(100 * v / tot) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v / tot >= 0.02}
    # fp64 thread-level instruction counts -> warp-level instructions per pair
    # fp64 thread-level instructions of the launch: (dadd + dmul + dfma per elapsed cycle) x elapsed cycles
    fp64_thread_inst = sum(f("smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op)
                           for op in ("dadd", "dmul", "dfma")) * f("smsp__cycles_elapsed.avg", f("sm__cycles_elapsed.avg"))
    rec = {
        "capture": os.path.basename(rep),
        "pairs_in_launch": pairs,
        "duration_us": round(f("gpu__time_duration.sum") * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0,
                                                             "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
                             .get(unit.get("gpu__time_duration.sum", "us"), 1.0), 2),
        "dram_bytes_read": scaled("dram__bytes_read.sum"),
        "dram_bytes_write": scaled("dram__bytes_write.sum"),
        "warp_instructions": f("smsp__inst_executed.sum"),
        "fp64_pipe_pct": round(f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), 1),
        "l1tex_data_pipe_pct": round(f("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"), 1),
        "smem_wavefronts": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "smem_bank_conflicts": f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "issue_active_pct": round(f("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
        "warps_active_per_sm": round(f("sm__warps_active.avg.per_cycle_active"), 1),
        "registers": int(f("launch__registers_per_thread")),
        "block": int(f("launch__block_size")), "grid": int(f("launch__grid_size")),
        "sm_cycles_active": {"min": f("sm__cycles_active.min"), "avg": f("sm__cycles_active.avg"),
                             "max": f("sm__cycles_active.max"), "elapsed": f("sm__cycles_elapsed.max")},
        "stalls_pct": stall_pct,
    }
    if fp64_thread_inst and pairs:
        rec["fp64_inst_per_pair"] = round(fp64_thread_inst / pairs)
        rec["fp64_mix_pct"] = {op: round(100 * f("smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op) /
                                         (fp64_thread_inst / f("smsp__cycles_elapsed.avg", f("sm__cycles_elapsed.avg"))))
                               for op in ("dadd", "dmul", "dfma")}
    # stall samples per code region (regions end at BAR / setmaxnreg / EXIT instructions)
    src = ncu_csv(rep, "source", kernel_filter)
    sh = src[1]
    ix = {h: i for i, h in enumerate(sh)}
    cols = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]

    def val(r, h):
        try:
            return float(r[ix[h]])
        except Exception:
            return 0.0

    regions, cur = [], {"samples": 0, "inst": 0, "stalls": {}}
    for r in src[2:]:
        s = r[ix["Source"]]
        cur["samples"] += val(r, "# Samples")
        cur["inst"] += val(r, "Instructions Executed")
        for h in cols:
            cur["stalls"][h[6:]] = cur["stalls"].get(h[6:], 0) + val(r, h)
        toks = s.split()
        op = (toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else ""))
        if op.startswith(("BAR", "USETMAXREG", "EXIT")):
            cur["end"] = " ".join(toks[:3])
            regions.append(cur)
            cur = {"samples": 0, "inst": 0, "stalls": {}}
    total = sum(c["samples"] for c in regions) or 1.0
    lines = []
    lines.append(f"kernel {kernel}   capture {os.path.basename(rep)}   pairs in launch {pairs}")
    for k, v in rec.items():
        if k not in ("capture", "pairs_in_launch"):
            lines.append(f"    {k:28s} {v}")
    lines.append("    stall samples per code region (region = instructions up to the named barrier):")
    for c in regions:
        if c["samples"] < 0.01 * total:
            continue
        top = sorted(c["stalls"].items(), key=lambda kv: -kv[1])[:6]
        lines.append("      %5.1f%% of samples, %10d warp-inst, ends at %-34s %s" % (
            100 * c["samples"] / total, c["inst"], c["end"], " ".join("%s=%d" % kv for kv in top)))
    with open(dst, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    d = load_summary()
    d["full_capture_" + kernel] = rec
    save_summary(d)
    print("\n".join(lines))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        cmd = sys.argv[4] if len(sys.argv) > 4 else \
            "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
        launches(sys.argv[2], sys.argv[3], cmd)
    else:
        full(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0,
             sys.argv[5] if len(sys.argv) > 5 else None)
