"""Join an ncu launch list (gpu__time_duration.sum, --csv) with the PC_GEMM_LOG=1 shape log of the same run:
per-GEMM-launch TFLOP/s and algorithmic GB/s, and the share of every kernel family.
  PC_GEMM_LOG=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv python tools/rn_driver.py 64 1 2> S.log
  python tools/rn_breakdown.py L.csv S.log [passes to skip]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
launches = []
for r in rows[hi + 1:]:
    if len(r) > mv:
        name = re.sub(r"\(.*", "", r[kn]).replace("void pc::<unnamed>::", "").replace("pc::<unnamed>::", "").replace("pc::", "")
        launches.append((name, float(r[mv].replace(",", "")) / 1e3))
shapes = [dict((k, int(v)) for k, v in re.findall(r"(\w+)=(\d+)", l)) for l in open(sys.argv[2], errors="ignore") if l.startswith("[gemm]")]
gemms = [i for i, l in enumerate(launches) if l[0].startswith("gemm_tn_kernel")]
assert len(gemms) == len(shapes), (len(gemms), len(shapes))
# the last encode_image pass: from the last stem_im2col on
start = max(i for i, l in enumerate(launches) if l[0].startswith("stem_im2col") or l[0].startswith("patchify"))
tot = collections.defaultdict(float)
T = 0.0
print(f"{'us':>8} {'TF/s':>7} {'GB/s':>7}  launch")
for i in range(start, len(launches)):
    name, us = launches[i]
    T += us
    if i in gemms:
        s = shapes[gemms.index(i)]
        taps = max(1, s["taps"])
        fl = 2.0 * s["M"] * s["N"] * s["K"] * taps
        by = 2.0 * (s["M"] * s["K"] + s["N"] * s["K"] * taps + s["M"] * s["N"] * (2 if s["res"] else 1))
        key = f"gemm M={s['M']} N={s['N']} K={s['K']}{' 3x3' if s['taps'] else ''}{' +res' if s['res'] else ''}"
        print(f"{us:8.1f} {fl / us / 1e6:7.0f} {by / us / 1e3:7.0f}  {key}")
        tot["gemm 3x3" if s["taps"] else ("gemm 1x1 +res" if s["res"] else "gemm 1x1")] += us
    else:
        print(f"{us:8.1f} {'':7} {'':7}  {name}")
        tot[name] += us
print(f"pass total {T:.0f} us")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"  {v:9.1f} us {100 * v / T:5.1f} %  {k}")
