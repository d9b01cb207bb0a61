"""Top stall lines of one captured launch: python scripts/ncu_hot.py <rep> <launch-id> [n]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:120])
hdr = rows[1]
i_src, i_smp, i_exec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = []
for r in rows[2:]:
    if len(r) > i_exec and r[i_smp].isdigit():
        body.append(r)
    elif r and r[0] == 'Kernel Name':
        break
tot = sum(int(r[i_smp] or 0) for r in body)
print("total samples", tot, "instructions", sum(int(r[i_exec] or 0) for r in body))
for idx, r in sorted(enumerate(body), key=lambda t: -int(t[1][i_smp] or 0))[:n]:
    print(f"{idx:5d} {int(r[i_smp]):7d} {100.0 * int(r[i_smp]) / max(tot, 1):5.1f}%  exec={r[i_exec]:>8}  {r[i_src].strip()[:90]}")
