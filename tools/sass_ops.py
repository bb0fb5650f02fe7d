"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md): UTC*MMA
(tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
MUFU, HMMA (legacy mma.sync: only the one-row helper warp of attention6's L = 257 variant may use it).   python tools/sass_ops.py [lib.so] > profiles/r02_sass_ops.txt"""
import re
import subprocess
import sys
from collections import OrderedDict

lib = sys.argv[1] if len(sys.argv) > 1 else "proto-clip_b200/libprotoclip_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = OrderedDict([("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("UTCHMMA.2CTA", r"UTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"),
                    ("STTM", r"\bSTTM"), ("UTMALDG", r"UTMALDG"), ("UTMASTG", r"UTMASTG"), ("UTCBAR", r"UTCBAR"),
                    ("SYNCS", r"\bSYNCS"), ("MUFU", r"\bMUFU"), ("HMMA", r"\bHMMA"), ("FFMA2", r"\bFFMA2")])
fn, rows = None, OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(CUtensorMap_st.*", "(...)", fn).replace("pc::(anonymous namespace)::", "")
        rows[fn] = OrderedDict((k, 0) for k in pats)
        rows[fn]["instrs"] = 0
        continue
    if fn and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        rows[fn]["instrs"] += 1
        for k, p in pats.items():
            if re.search(p, line):
                rows[fn][k] += 1
print(f"# {lib}: SASS mnemonic counts per kernel (cuobjdump -sass); sm_100a only")
hdr = list(pats) + ["instrs"]
print(f"{'kernel':78s} " + " ".join(f"{h:>8s}" for h in hdr))
tot = OrderedDict((h, 0) for h in hdr)
for fn, r in rows.items():
    print(f"{fn[:78]:78s} " + " ".join(f"{r[h]:8d}" for h in hdr))
    for h in hdr:
        tot[h] += r[h]
print(f"{'TOTAL':78s} " + " ".join(f"{tot[h]:8d}" for h in hdr))
