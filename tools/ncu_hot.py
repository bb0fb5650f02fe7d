"""Hot SASS lines of one kernel in an .ncu-rep: python tools/ncu_hot.py rep.ncu-rep [kernel-index] [top-n]"""
import csv
import io
import subprocess
import sys


def main(path, kidx=0, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    # split per kernel
    blocks, cur = [], []
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            if cur:
                blocks.append(cur)
            cur = [line]
        elif cur:
            cur.append(line)
    if cur:
        blocks.append(cur)
    b = blocks[kidx]
    print(b[0][:160])
    rows = list(csv.reader(io.StringIO("\n".join(b[1:]))))
    hdr = rows[0]
    si, ns = hdr.index("Source"), hdr.index("# Samples")
    ie = hdr.index("Instructions Executed")
    stall_cols = [j for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for i, r in enumerate(rows[1:]):
        try:
            data.append((int(r[ns]), i, r))
        except (ValueError, IndexError):
            pass
    total = sum(d[0] for d in data)
    print(f"total samples {total}, instructions {len(data)}")
    # stall totals
    tot = {hdr[j]: 0 for j in stall_cols}
    for n, i, r in data:
        for j in stall_cols:
            tot[hdr[j]] += int(r[j] or 0)
    print("stall totals:", {k: v for k, v in sorted(tot.items(), key=lambda x: -x[1]) if v})
    # opcode histogram by executed instructions
    ops = {}
    for n, i, r in data:
        op = r[si].split()[0] if not r[si].strip().startswith("@") else r[si].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[ie] or 0)
    print("executed by opcode:", sorted(ops.items(), key=lambda x: -x[1])[:25])
    for n, i, r in sorted(data, reverse=True)[:top]:
        st = {hdr[j].replace("stall_", ""): int(r[j] or 0) for j in stall_cols if int(r[j] or 0)}
        st = dict(sorted(st.items(), key=lambda x: -x[1])[:3])
        print(f"{n:6d} {100 * n / total:5.1f}%  #{i:4d} exec={r[ie]:>8s} {r[si].strip()[:70]:70s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
