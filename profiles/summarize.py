#!/usr/bin/env python
"""Turns the ncu outputs of profiles/run_ncu.sh (in gpurun_out/) into the committed summaries
under profiles/:  python profiles/summarize.py r01b

  <tag>_launches.csv     the raw ncu launch list (gpu__time_duration per launch)
  <tag>_summary.md       per-kernel share of the step + key metrics of every --set full capture
                         (duration, IPC, issue-slot / FMA-pipe utilisation, warp execution
                         efficiency, occupancy, L1/L2 hit rates, DRAM bytes, top stall reasons,
                         hottest source lines)
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.
"""
import collections
import csv
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
HERE = os.path.join(ROOT, "profiles")

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC (per SM, active)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % of peak"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe % of peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe % of peak"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "avg active threads / warp instr"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("l1tex__t_sector_hit_rate.pct", "L1/TEX hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
]


def launch_summary(tag, lines):
    path = os.path.join(OUT, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    shutil.copy(path, os.path.join(HERE, f"{tag}_launches.csv"))
    rows = csv.DictReader([l for l in open(path) if not l.startswith("==")])
    agg, total = collections.OrderedDict(), 0.0
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        ms = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r["Metric Unit"], 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
        total += ms
    lines.append(f"## Launch list (`launches_{tag}.csv`, whole `bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs` run)\n")
    lines.append("The run contains the grid-accelerated arm (value + e2e), the brute-force roofline arm and the\n"
                 "FMA-peak microbenchmarks; shares are of the summed kernel time of the whole run.\n")
    lines.append("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {ms:.3f} | {100 * ms / total:.1f} % |")
    lines.append("")


def grid_counters(tag, lines):
    """Instruction counters of every k_trace_grid launch of two C4 frames -> <tag>_grid_counters.json: the numerator of the
    lane-issue roofline bench.py reports (thread instructions per ray) and the IPC x lanes product under ncu."""
    import json
    path = os.path.join(OUT, f"gridcounters_{tag}.csv")
    if not os.path.exists(path):
        return
    shutil.copy(path, os.path.join(HERE, f"{tag}_gridcounters.csv"))
    rows = list(csv.DictReader([l for l in open(path) if not l.startswith("==")]))
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"])})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        d[r["Metric Name"] + "/unit"] = r["Metric Unit"]
    launches = list(per.values())
    half = len(launches) // 2          # two identical frames were captured: use the second (warm)
    frame = launches[half:] if half else launches
    to_ns = lambda d: d["gpu__time_duration.sum"] * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(d.get("gpu__time_duration.sum/unit", "ns"), 1.0)
    thread_inst = sum(d["smsp__thread_inst_executed.sum"] for d in frame)
    warp_inst = sum(d["smsp__inst_executed.sum"] for d in frame)
    ns = sum(to_ns(d) for d in frame)
    rays = 111009057   # C4 at 3840x2160: every Scene::trace call of the frame goes through one of these launches
    peak_per_ns = 148 * 4 * 32 * 1.965   # thread instructions per ns at 1965 MHz
    out = {"workload": "C4 3840x2160 (tools/prof_grid.py, host-sized loop: every launch exactly sized)", "launches": len(frame),
           "rays": rays, "thread_inst": thread_inst, "warp_inst": warp_inst, "thread_inst_per_ray": thread_inst / rays,
           "warp_inst_per_ray": warp_inst / rays, "avg_active_lanes": thread_inst / warp_inst,
           "sum_kernel_ms_under_ncu": ns * 1e-6, "ipc_lanes_product": thread_inst / (ns * peak_per_ns),
           "note": "ipc_lanes_product = thread instructions / (kernel time under ncu x 148 SM x 4 x 32 x 1.965 GHz) = IPC/4 x lanes/32, "
                   "serialised and cold-cache; bench.py multiplies thread_inst_per_ray by the rays/s it measures"}
    with open(os.path.join(HERE, f"{tag}_grid_counters.json"), "w") as f:
        json.dump(out, f, indent=1)
    lines.append(f"## Instruction counters of `k_trace_grid` over one C4 frame (`{tag}_gridcounters.csv`)\n")
    lines.append(f"- launches {len(frame)}, rays {rays}: **{out['thread_inst_per_ray']:.0f} thread instructions per ray**, "
                 f"{out['warp_inst_per_ray']:.1f} warp instructions per ray, {out['avg_active_lanes']:.1f} of 32 lanes active on average")
    lines.append(f"- lane-issue utilisation under ncu (IPC/4 x lanes/32): **{out['ipc_lanes_product']:.3f}** "
                 f"(sum of kernel times {out['sum_kernel_ms_under_ncu']:.2f} ms, serialised)")
    lines.append("| launch | kernel | ms | warp instr | lanes | IPC |\n|---:|---|---:|---:|---:|---:|")
    for k, d in enumerate(frame):
        lines.append(f"| {k} | `{d['name']}` | {to_ns(d) * 1e-6:.3f} | {d['smsp__inst_executed.sum']:.3g} | "
                     f"{d.get('smsp__thread_inst_executed_per_inst_executed.ratio', 0):.1f} | {d.get('sm__inst_executed.avg.per_cycle_active', 0):.2f} |")
    lines.append("")


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        d = {"name": r[ix["Kernel Name"]], "grid": r[ix.get("Grid Size", 0)] if "Grid Size" in ix else ""}
        for key, label in KEYS:
            if key in ix:
                d[label] = f"{r[ix[key]]} {units[ix[key]]}".strip()
        stalls = []
        for h, i in ix.items():
            m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active\.ratio", h)
            if m and "not_issued" not in h:
                try:
                    stalls.append((float(r[i]), m.group(1)))
                except ValueError:
                    pass
        d["stalls"] = ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:6])
        res.append(d)
    return res


def hot_lines(rep, top=8):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    agg, cur_file, cur_fn = {}, None, None
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = os.path.basename(r[1]); continue
        if r[0] == "Function Name":
            cur_fn = re.sub(r"\(.*", "", r[1]).replace("void rg::", ""); continue
        if r[0].isdigit() and len(r) > 8:
            s = int(r[6]) if r[6].isdigit() else 0
            ex = int(r[7]) if r[7].isdigit() else 0
            th = int(r[8]) if r[8].isdigit() else 0
            a = agg.setdefault((cur_fn, cur_file, int(r[0]), r[1].strip()[:80]), [0, 0, 0])
            a[0] += s; a[1] += ex; a[2] += th
    per_fn = collections.defaultdict(list)
    for (fn, f, ln, src), (s, ex, th) in agg.items():
        per_fn[fn].append((s, f, ln, src, ex, th))
    res = {}
    for fn, items in per_fn.items():
        tot = sum(i[0] for i in items) or 1
        res[fn] = [(100.0 * s / tot, f, ln, src, th / max(ex, 1)) for s, f, ln, src, ex, th in sorted(items, reverse=True)[:top]]
    return res


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    lines = [f"# ncu summary `{tag}` (B200, `--clock-control none`)\n",
             "Generated by `profiles/summarize.py` from the outputs of `profiles/run_ncu.sh` (which first runs the\n"
             "same command without ncu and requires exit 0).  Absolute times under ncu are cold-cache and\n"
             "serialised; bench.py's CUDA-event numbers are the timings of record.\n"]
    launch_summary(tag, lines)
    grid_counters(tag, lines)
    for rep in sorted(f for f in os.listdir(OUT) if f.endswith(f"_{tag}.ncu-rep")):
        path = os.path.join(OUT, rep)
        lines.append(f"## `--set full` capture `{rep}`\n")
        hots = hot_lines(path)
        for d in raw_metrics(path):
            lines.append(f"### `{d['name'][:70]}`  grid {d['grid']}\n")
            for _, label in KEYS:
                if label in d:
                    lines.append(f"- {label}: {d[label]}")
            lines.append(f"- top stall reasons (warps per issue-active cycle): {d['stalls']}")
            lines.append("")
        for fn, items in hots.items():
            lines.append(f"Hottest source lines of `{fn}` (share of stall samples, avg active threads):\n")
            for share, f, ln, src, thr in items:
                lines.append(f"- {share:4.1f} %  `{f}:{ln}`  thr {thr:4.1f}  `{src}`")
            lines.append("")
    with open(os.path.join(HERE, f"{tag}_summary.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", os.path.join(HERE, f"{tag}_summary.md"))


if __name__ == "__main__":
    main()
