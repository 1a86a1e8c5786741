"""SASS opcode census of libstito.so per kernel (developer tool): proves which kernels use tcgen05 (UTCHMMA / UTCBAR /
LDTM), TMA (UTMALDG / UBLKCP), thread-block clusters / DSMEM and mbarriers.   python scripts/sass_census.py > profiles/<tag>_sass_census.csv"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "st_ito_b200", "libstito.so")
sass = subprocess.check_output(["cuobjdump", "-sass", lib], text=True)
ops = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMAPF", "UBLKCP", "SYNCS", "UCGABAR", "ST.E", "LD.E", "LDS", "STS", "DFMA", "DMUL", "DADD",
       "FFMA", "MUFU", "SHFL", "BAR", "ATOM", "RED", "MAPA", "LDG", "STG", "HMMA", "NANOSLEEP"]
counts, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.check_output(["c++filt", m.group(1)], text=True).strip()
        name = re.sub(r"stito::\(anonymous namespace\)::", "", name).split("(")[0].replace("void ", "")
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        counts[name]["_total"] += 1
        for o in ops:
            if op == o or op.startswith(o + ".") or (o in ("ST.E", "LD.E") and op.startswith(o)):
                counts[name][o] += 1
print("# SASS opcode census of st_ito_b200/libstito.so (cuobjdump -sass, sm_100a). UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, "
      "LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk (shared -> DSMEM), SYNCS = mbarrier ops, UCGABAR = cluster barrier, MAPA = DSMEM address map")
print("kernel,instructions," + ",".join(ops))
for k, c in counts.items():
    print(f"\"{k}\",{c['_total']}," + ",".join(str(c[o]) for o in ops))
