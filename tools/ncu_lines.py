"""Per-CUDA-source-line instruction and stall-sample shares of one kernel in an .ncu-rep (needs -lineinfo)."""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
h = rows[hi]
ie, isamp = h.index("Instructions Executed"), h.index("# Samples")
data = {}
for r in rows[hi + 1:]:
    if len(r) <= max(ie, isamp) or not r[0].isdigit() or r[2] != "-":
        continue  # keep the per-line aggregate rows (Address == "-")
    ln = int(r[0])
    if ln in data:
        continue
    data[ln] = (r[1], int(r[ie] or 0), int(r[isamp] or 0))
tot = sum(d[1] for d in data.values()); ts = sum(d[2] for d in data.values())
print("kernel", kern, "warp instructions", tot, "samples", ts)
for ln, d in sorted(data.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5d  instr %5.1f%%  samples %5.1f%%  %s" % (ln, 100.0 * d[1] / max(tot, 1), 100.0 * d[2] / max(ts, 1), d[0].strip()[:120]))
