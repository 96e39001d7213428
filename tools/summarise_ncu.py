#!/usr/bin/env python
"""Turn the raw ncu artefacts that gpurun brings back (gpurun_out/) into the tracked summaries under
profiles/:  a per-kernel launch list of one bench step, the key metrics of the ray-score kernels from the
`--set full` capture, and profiles/ncu_traffic.json (DRAM bytes per launch) which bench.py reports as
`roofline.traffic`.

  python tools/summarise_ncu.py --round r1 --launches gpurun_out/launches_r1.csv --rep gpurun_out/prof_score_r1.ncu-rep
"""
import argparse
import collections
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__inst_executed_pipe_xu", "smsp__inst_executed_pipe_xu.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "smsp__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "smsp__inst_executed.sum", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(path, out_md, title):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit.startswith("us") else v * 1e3)
        a = agg.setdefault(r["Kernel Name"], [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += v_us
    tot = sum(a[1] for a in agg.values())
    with open(out_md, "w") as f:
        f.write(f"# {title}\n\nSource: `{os.path.relpath(path, ROOT)}` (ncu `--metrics gpu__time_duration.sum --clock-control none`, "
                "profiling started at the timed region). Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n\n")
        f.write(f"{len(rows)} launches, {tot / 1e3:.3f} ms of device time in total.\n\n| share | total us | launches | us/launch | grid | block | kernel |\n|---|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {100 * a[1] / tot:5.1f}% | {a[1]:10.1f} | {a[0]} | {a[1] / a[0]:9.1f} | {a[2]} | {a[3]} | `{k[:110]}` |\n")
    return agg, tot


def rep_metrics(rep, out_md, out_json, title, note=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    traffic = {}
    with open(out_md, "w") as f:
        f.write(f"# {title}\n\nSource: `{os.path.relpath(rep, ROOT)}` (`ncu --set full --clock-control none --import-source on`), read with "
                "`ncu -i ... --page raw --csv`.\n\n")
        for d in data:
            name = d[col["Kernel Name"]]
            f.write(f"## `{name[:100]}` (launch id {d[col['ID']]})\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in col:
                    f.write(f"| {k} | {d[col[k]]} | {units[col[k]]} |\n")
            try:
                rd = float(d[col["dram__bytes_read.sum"]].replace(",", ""))
                wr = float(d[col["dram__bytes_write.sum"]].replace(",", ""))
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd *= scale.get(units[col["dram__bytes_read.sum"]], 1)
                wr *= scale.get(units[col["dram__bytes_write.sum"]], 1)
                key = "score_pass1" if "ILi1E" in name or "<1>" in name or "<(int)1>" in name else "score_pass2"
                traffic[key + "_profiled"] = {"dram_bytes": rd + wr, "kernel": name[:80]}
                f.write(f"\nDRAM traffic = {rd + wr:.4g} B per launch.\n\n")
            except (KeyError, ValueError):
                pass
    # speed-of-light / scheduler / occupancy rows of the details page
    det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    drows = list(csv.reader(io.StringIO(det)))
    if drows:
        h = drows[0]
        want = ("SM Frequency", "DRAM Frequency", "Duration", "Memory Throughput", "DRAM Throughput", "L2 Cache Throughput",
                "Compute (SM) Throughput", "Executed Ipc Active", "Issue Slots Busy", "Mem Busy", "Mem Pipes Busy", "L2 Hit Rate",
                "No Eligible", "Eligible Warps Per Scheduler", "Warp Cycles Per Issued Instruction", "Theoretical Occupancy",
                "Achieved Occupancy", "Registers Per Thread", "Dynamic Shared Memory Per Block", "Cluster Size")
        with open(out_md, "a") as f:
            f.write("## ncu details page (speed of light, scheduler, occupancy)\n\n| kernel | section | metric | value | unit |\n|---|---|---|---|---|\n")
            for r in drows[1:]:
                d = dict(zip(h, r))
                if d.get("Metric Name") in want:
                    f.write(f"| `{d['Kernel Name'][:32]}` | {d['Section Name']} | {d['Metric Name']} | {d['Metric Value']} | {d['Metric Unit']} |\n")
            if note:
                f.write("\n" + note + "\n")
    json.dump(traffic, open(out_json, "w"), indent=1)
    return traffic


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r1")
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--name", default="score_kernels", help="basename of the summary written under profiles/")
    ap.add_argument("--note", default="", help="a 'Reading:' paragraph appended to the summary")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if a.launches:
        launches(a.launches, os.path.join(ROOT, "profiles", f"launches_{a.round}.md"), f"Launch list of the timed bench steps ({a.round})")
    if a.rep:
        print(rep_metrics(a.rep, os.path.join(ROOT, "profiles", f"{a.name}_{a.round}.md"),
                          os.path.join(ROOT, "profiles", f"ncu_profiled_{a.name}_{a.round}.json"),
                          f"{a.name}, ncu --set full ({a.round})", a.note))
