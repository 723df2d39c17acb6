"""Summarise ncu outputs into small text files under profiles/ (the .ncu-rep itself stays in gpurun_out/).

    python tools/ncu_summary.py launches gpurun_out/r01_launches.csv > profiles/r01_launches_summary.txt
    python tools/ncu_summary.py full gpurun_out/r01_bwd_v4.ncu-rep  > profiles/r01_bwd_full_summary.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warps_eligible.avg.per_cycle_active",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] == "ns" else (v * 1000.0 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    mine = {k: v for k, v in agg.items() if "eslam::" in k}
    tot = sum(a[1] for a in mine.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv  python bench.py --steps 2 --warmup 1 --profile-only")
    print("# (the bench's own launch sequence: device-resident mapping calls, the backward kernel alone, tracking, the e2e")
    print("#  drop-in calls; cold-cache and serialised under ncu: compare SHARES, not absolutes).  Only this repo's kernels")
    print("#  are listed one by one; torch's kernels (scene synthesis, RNG draws, fills, copies) are summed in the last line.")
    print(f"{'us total':>10} {'launches':>8} {'us/launch':>10} {'share':>7}  kernel")
    for n, (c, t) in sorted(mine.items(), key=lambda x: -x[1][1]):
        print(f"{t:10.1f} {c:8d} {t / c:10.1f} {100 * t / tot:6.1f}%  {n}")
    print(f"{tot:10.1f} {sum(a[0] for a in mine.values()):8d}  (all eslam kernels)")
    other = sum(a[1] for k, a in agg.items() if "eslam::" not in k)
    print(f"{other:10.1f} {sum(a[0] for k, a in agg.items() if 'eslam::' not in k):8d}  (torch kernels: scene synthesis, RNG draws, fills)")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none --import-source on  ({path})")
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:72s} {r[i]:>16s} {units[i]}")
        print("  stall reasons (warps stalled per issue-active cycle):")
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                v = float(r[i])
                if v > 0.05:
                    print(f"    {h.split('issue_stalled_')[1].split('_per')[0]:24s} {v:.2f}")


def to_json(path):
    """Per kernel (first launch of each name): duration and DRAM bytes per launch, as JSON (bench.py reads
    profiles/r02_ncu_summary.json for roofline.traffic)."""
    import json

    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]

    def val(r, k):
        i = hdr.index(k)
        v = float(r[i].replace(",", ""))
        u = units[i]
        return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)

    res = {"source": path, "how": "ncu --set full --clock-control none (cold cache, serialised)"}
    for r in rows[2:]:
        name = re.sub(r"<.*|\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "").replace("eslam::", "")
        if name in res:
            continue
        res[name] = {"dram_bytes_per_launch": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
                     "duration_us": float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")),
                     "l2_read_sectors": float(r[hdr.index("lts__t_sectors_srcunit_tex_op_read.sum")].replace(",", "")),
                     "l2_red_sectors": float(r[hdr.index("lts__t_sectors_srcunit_tex_op_red.sum")].replace(",", "")),
                     "warps_active_pct": float(r[hdr.index("sm__warps_active.avg.pct_of_peak_sustained_active")]),
                     "issue_active_pct": float(r[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")])}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "json": to_json}[sys.argv[1]](sys.argv[2])
