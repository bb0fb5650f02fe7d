"""One line per kernel of an .ncu-rep: duration, DRAM bytes, achieved HBM GB/s vs the measured copy peak, tensor / XU
pipe utilisation, registers.   python tools/ncu_table.py rep.ncu-rep [peak_gbs]"""
import csv
import io
import json
import os
import subprocess
import sys

path = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
if peak is None:
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale_to=None):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return float("nan")
    v = float(r[i].replace(",", ""))
    u = units[i]
    if scale_to == "bytes":
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if scale_to == "us":
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
    return v


print(f"# {os.path.basename(path)}: ncu --set full --clock-control none; HBM peak = {peak:.0f} GB/s (MEASURED_PEAKS.json copy rate)")
print(f"{'kernel':58s} {'grid':>6s} {'regs':>4s} {'us':>8s} {'dram rd MB':>10s} {'dram wr MB':>10s} {'GB/s':>7s} {'of peak':>7s} {'tensor%':>7s} {'xu%':>6s} {'issue%':>6s}")
for r in rows[2:]:
    name = r[col["Kernel Name"]].replace("void ", "").replace("pc::<unnamed>::", "").replace("<unnamed>::", "")
    name = name.split("(CUtensorMap")[0].split("(const")[0][:58]
    us = val(r, "gpu__time_duration.sum", "us")
    rd, wr = val(r, "dram__bytes_read.sum", "bytes"), val(r, "dram__bytes_write.sum", "bytes")
    gbs = (rd + wr) / us / 1e3 if us > 0 else float("nan")
    print(f"{name:58s} {int(val(r, 'launch__grid_size')):6d} {int(val(r, 'launch__registers_per_thread')):4d} {us:8.2f} "
          f"{rd / 1e6:10.2f} {wr / 1e6:10.2f} {gbs:7.0f} {gbs / peak:7.3f} "
          f"{val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):7.1f} "
          f"{val(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f}")
