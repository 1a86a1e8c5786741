"""Summarise an `ncu --set full` raw CSV of the HBM-bound kernels (DSP chain, log-mel, conv1, conv reverb) into a tracked
table under profiles/ (developer tool):  duration, DRAM bytes, achieved GB/s and fraction of the measured HBM peak,
issue-slot utilisation, occupancy, and the three largest warp-stall reasons (pc sampling).

    python scripts/make_hbm_summary.py <raw.csv> <out.csv> "<header comment>"
"""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, out, comment = sys.argv[1], sys.argv[2], sys.argv[3]
peak = 6549.4
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.isfile(pk):
    peak = json.load(open(pk))["hbm_gbs"]
rr = list(csv.reader(open(raw)))
hdr, units = rr[0], rr[1]
col = {h: i for i, h in enumerate(hdr)}
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
tmult = {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
stall_cols = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
rows = []
for r in rr[2:]:
    def val(name):
        return float(r[col[name]].replace(",", "")) if r[col[name]] not in ("", "n/a") else 0.0
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("stito::(anonymous namespace)::", "")
    us = val("gpu__time_duration.sum") * tmult[units[col["gpu__time_duration.sum"]]]
    rd = val("dram__bytes_read.sum") * mult[units[col["dram__bytes_read.sum"]]]
    wr = val("dram__bytes_write.sum") * mult[units[col["dram__bytes_write.sum"]]]
    stalls = sorted(((val(h), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h in stall_cols), reverse=True)
    tot = sum(s[0] for s in stalls) or 1.0
    top = "; ".join(f"{n} {100 * v / tot:.0f}%" for v, n in stalls[:3])
    rows.append([name, r[col["launch__grid_size"]], r[col["launch__block_size"]], f"{us:.1f}", f"{rd / 1e6:.1f}", f"{wr / 1e6:.1f}",
                 f"{(rd + wr) / us / 1e3:.0f}", f"{(rd + wr) / us / 1e3 / peak:.3f}",
                 r[col["sm__issue_active.avg.pct_of_peak_sustained_elapsed"]],
                 r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]],
                 r[col["launch__registers_per_thread"]], r[col["launch__shared_mem_per_block_dynamic"]], top])
with open(out, "w") as f:
    f.write(f"# {comment}\n")
    w = csv.writer(f)
    w.writerow(["kernel", "grid", "block", "us", "dram_read_MB", "dram_write_MB", "achieved_GBps", f"frac_of_hbm_peak_{peak:.0f}GBps",
                "issue_active_pct", "warps_active_pct", "regs", "dyn_smem_B", "top_stalls_pcsamp"])
    w.writerows(rows)
print(open(out).read())
