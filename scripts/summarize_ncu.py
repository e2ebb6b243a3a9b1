"""Summarise gpurun_out/<tag>_full.ncu-rep and <tag>_launches.csv into profiles/ (tracked).

  python scripts/summarize_ncu.py r01
writes profiles/<tag>_ncu_summary.json, profiles/<tag>_ncu_raw_selected.csv,
profiles/<tag>_launches.csv (trimmed launch list) and profiles/<tag>_top_stalls.txt.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]


def short(name):
    for k in ("k_fused_tc", "k_blend_tc", "k_lbs_tc", "k_pose_chain", "k_blend_fma", "k_lbs_fma", "k_regress_joints", "k_pack"):
        if k in name:
            return k
    return name[:40]


def main(tag):
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_full.ncu-rep")
    out_dir = os.path.join(ROOT, "profiles")
    os.makedirs(out_dir, exist_ok=True)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    rep2 = os.path.join(ROOT, "gpurun_out", f"{tag}_full_unfused.ncu-rep")       # second capture: the unfused kernels
    if os.path.exists(rep2):
        raw2 = subprocess.run(["ncu", "-i", rep2, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows2 = list(csv.reader(io.StringIO(raw2)))
        if rows2 and rows2[0] == hdr:
            data = data + rows2[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    summary = {"tag": tag, "source": f"ncu --set full --clock-control none ({os.path.basename(rep)})", "kernels": {}}
    sel = [["kernel"] + KEEP, ["unit"] + [units[idx[k]] if k in idx else "" for k in KEEP]]
    for r in data:
        name = short(r[idx["Kernel Name"]])
        vals = {k: r[idx[k]] for k in KEEP if k in idx}
        sel.append([name] + [vals.get(k, "") for k in KEEP])
        f = lambda k: float(vals[k]) if vals.get(k) not in (None, "") else None
        to_bytes = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = f("dram__bytes_read.sum") * to_bytes.get(units[idx["dram__bytes_read.sum"]], 1)
        wr = f("dram__bytes_write.sum") * to_bytes.get(units[idx["dram__bytes_write.sum"]], 1)
        summary["kernels"][name] = {
            "duration_us_under_ncu": f("gpu__time_duration.sum"),
            "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
            "dram_throughput_pct_of_ncu_peak": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "lts_throughput_pct": f("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            "tensor_pipe_active_pct": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": f("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
            "registers_per_thread": f("launch__registers_per_thread"), "grid": f("launch__grid_size"),
            "block": f("launch__block_size"), "waves_per_sm": f("launch__waves_per_multiprocessor"),
        }
    with open(os.path.join(out_dir, f"{tag}_ncu_summary.json"), "w") as fo:
        json.dump(summary, fo, indent=1)
    with open(os.path.join(out_dir, f"{tag}_ncu_raw_selected.csv"), "w", newline="") as fo:
        csv.writer(fo).writerows(sel)
    # launch list: keep kernel name + duration only
    ll = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
    if os.path.exists(ll):
        lines = [l for l in open(ll) if l.startswith('"')]
        rows = list(csv.reader(lines))
        h = rows[0]
        ki, vi, mi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
        with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w", newline="") as fo:
            w = csv.writer(fo)
            w.writerow(["launch", "kernel", "gpu__time_duration.sum (ns, cold-cache, serialised)"])
            agg = {}
            for i, r in enumerate(rows[1:]):
                if r[mi] != "gpu__time_duration.sum":
                    continue
                w.writerow([i, short(r[ki]), r[vi]])
                agg.setdefault(short(r[ki]), []).append(float(r[vi].replace(",", "")))
            w.writerow([])
            w.writerow(["kernel", "launches", "mean_ns"])
            for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
                w.writerow([k, len(v), round(sum(v) / len(v), 1)])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    tmp = os.path.join(ROOT, "gpurun_out", f"{tag}_source.csv")
    open(tmp, "w").write(src)
    top = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_top_stalls.py"), tmp, "14"],
                         capture_output=True, text=True).stdout
    open(os.path.join(out_dir, f"{tag}_top_stalls.txt"), "w").write(top)
    print(json.dumps(summary, indent=1)[:1500])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
