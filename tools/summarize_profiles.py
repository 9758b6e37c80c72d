"""Turn gpurun_out/<tag>_full.ncu-rep and <tag>_launches.csv into the tracked summaries under profiles/.
usage: python tools/summarize_profiles.py <tag>"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active"]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    tag = sys.argv[1]
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_full.ncu-rep")
    out_dir = os.path.join(ROOT, "profiles")
    os.makedirs(out_dir, exist_ok=True)
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    lines, traffic = [], {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines.append(f"== {name}")
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"   {k:78s} {r[i]:>18s} {units[i]}")
        def num(k):
            i = hdr.index(k)
            v = float(r[i])
            u = units[i].lower()
            return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}.get(u, 1)
        tot = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        lines.append(f"   dram traffic per launch (read+write)                                          {tot / 1e9:18.4f} GB")
        key = "histogram" if "histogram" in name else "onesweep"
        traffic.setdefault(key, tot)
    open(os.path.join(out_dir, f"{tag}_ncu_full_summary.txt"), "w").write("\n".join(lines) + "\n")
    # stall reasons of the onesweep kernel
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:onesweep"],
                         capture_output=True, text=True).stdout
    tmp = os.path.join(ROOT, "gpurun_out", f"{tag}_src.csv")
    open(tmp, "w").write(src)
    top = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_top.py"), tmp, "30"], capture_output=True, text=True).stdout
    open(os.path.join(out_dir, f"{tag}_onesweep_stalls.txt"), "w").write(top)
    # launch list: per-kernel totals and share of the step
    lcsv = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
    if os.path.exists(lcsv):
        text = [l for l in open(lcsv) if l.startswith('"')]
        rd = list(csv.reader(text))
        h = rd[0]
        ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
        agg = {}
        for r in rd[1:]:
            v = float(r[vi].replace(",", ""))
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
            short = r[ki].split("(")[0].replace("void ", "")[:70]
            a = agg.setdefault(short, [0, 0.0])
            a[0] += 1
            a[1] += v
        total = sum(a[1] for a in agg.values())
        with open(os.path.join(out_dir, f"{tag}_launch_list_summary.txt"), "w") as f:
            f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {len(rd) - 1} launches, {total:.3f} ms\n")
            for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"  {k:72s} launches={c:4d} total={ms:9.3f} ms share={100 * ms / total:5.1f}% avg={ms / c:8.4f} ms\n")
        import shutil
        shutil.copy(lcsv, os.path.join(out_dir, f"{tag}_launches.csv"))
    json.dump({"onesweep_u32_keys_2^28_dram_bytes_per_launch": traffic.get("onesweep"),
               "histogram_u32_2^28_dram_bytes_per_launch": traffic.get("histogram"), "source": f"profiles/{tag}_ncu_full_summary.txt"},
              open(os.path.join(out_dir, "traffic.json"), "w"), indent=1)
    print(open(os.path.join(out_dir, f"{tag}_ncu_full_summary.txt")).read()[:3000])


main()
