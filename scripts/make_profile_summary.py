"""Turn the ncu captures of a gpurun call into the tracked summaries under profiles/ (developer tool).

    python scripts/make_profile_summary.py <tag> <launches.csv> <conv_full.ncu-rep> [n_last_launches]

Writes profiles/<tag>_generation_kernels.csv (per-kernel device time of the last generation),
profiles/<tag>_conv_ncu_full_summary.csv (selected metrics of every conv launch) and
profiles/<tag>_conv_traffic.json (dram bytes per generation of the conv stack; bench.py reports it as
roofline.traffic)."""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
n_last = int(sys.argv[4]) if len(sys.argv) > 4 else 24
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)

rows = list(csv.DictReader([l for l in open(launches) if not l.startswith("==")]))
out = [f"# {tag}: per-kernel device time of ONE generation (P=64, 10 s stereo, EQ+Comp+Reverb); ncu --metrics "
       "gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)",
       "kernel,grid,block,us"]
tot = 0.0
for r in rows[-n_last:]:
    us = float(r["Metric Value"]) / 1e3
    tot += us
    name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("unnamed>::", "")
    out.append(f"{name},{r['Grid Size'].replace(',', ' ')},{r['Block Size'].replace(',', ' ')},{us:.1f}")
out.append(f"TOTAL,,,{tot:.1f}")
open(os.path.join(ROOT, "profiles", f"{tag}_generation_kernels.csv"), "w").write("\n".join(out) + "\n")

raw = open(rep).read() if rep.endswith(".csv") else subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
want = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic"]
idx = [hdr.index(w) for w in want if w in hdr]
with open(os.path.join(ROOT, "profiles", f"{tag}_conv_ncu_full_summary.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow([f"# {tag}: ncu --set full --clock-control none -k regex:conv3x3|wino (the conv-stack launches of one "
                "generation, P=64, 10 s stereo); second row = units"])
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in rr[2:]:
        w.writerow([r[i] for i in idx])

def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * m[unit]

ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
it = hdr.index("gpu__time_duration.sum")
per = [{"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("unnamed>::", ""),
        "dram_read_bytes": to_bytes(r[ir], units[ir]), "dram_write_bytes": to_bytes(r[iw], units[iw]),
        "ms": float(r[it]) * {"ms": 1, "us": 1e-3, "s": 1e3, "usecond": 1e-3, "msecond": 1, "second": 1e3, "ns": 1e-6,
                              "nsecond": 1e-6}[units[it]]} for r in rr[2:]]
summary = {"tag": tag, "launches": len(per), "per_launch": per,
           "dram_bytes_per_generation": sum(p["dram_read_bytes"] + p["dram_write_bytes"] for p in per),
           "how": "dram__bytes_read.sum + dram__bytes_write.sum summed over the conv launches of one generation, "
                  "ncu --set full --clock-control none"}
json.dump(summary, open(os.path.join(ROOT, "profiles", f"{tag}_conv_traffic.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in summary.items() if k != "per_launch"}))
