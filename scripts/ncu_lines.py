"""Hottest source lines of ONE kernel of an .ncu-rep:  python scripts/ncu_lines.py rep kernel-regex [n]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
for i, r in enumerate(rows[:12]):
    if "Source" in r and "Instructions Executed" in r:
        hdr, start = r, i + 1
        break
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
lines, tot, totS = [], 0, 0
for r in rows[start:]:
    if len(r) > iI and r[0] != "" and r[2] == "-":
        try:
            n, s = int(r[iI]), int(r[iS])
        except ValueError:
            continue
        lines.append((n, s, r[0], r[1])); tot += n; totS += s
print(f"total inst {tot}, samples {totS}")
key = (lambda x: -x[0]) if len(sys.argv) > 4 and sys.argv[4] == "inst" else (lambda x: -x[1])
for n, s, ln, code in sorted(lines, key=key)[:nl]:
    print(f"  {s / max(totS, 1) * 100:5.1f}% smp {n / max(tot, 1) * 100:5.1f}% inst  L{ln}: {code.strip()[:110]}")
