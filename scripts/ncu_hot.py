"""Hot SASS regions of one kernel from an ncu report's source page.
    python scripts/ncu_hot.py <rep> [top]  -> instructions executed / stall samples per SASS line, grouped in runs."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = rows[2:]
tot_ex = sum(int(r[iex]) for r in body)
tot_s = sum(int(r[isamp]) for r in body)
print(f"total warp instructions executed {tot_ex:,}  samples {tot_s:,}  lines {len(body)}")
# print every line with its share, only the top-N by executed count plus context marks
thr = sorted((int(r[iex]) for r in body), reverse=True)[min(top, len(body) - 1)]
for k, r in enumerate(body):
    ex, s = int(r[iex]), int(r[isamp])
    if ex >= thr:
        print(f"{k:5d} {100.0 * ex / tot_ex:5.2f}% ex {100.0 * s / max(tot_s, 1):5.2f}% smp  {r[isrc].strip()}")
print("--- blocks of 20 lines: first line, %executed, %samples")
for b in range(0, len(body), 20):
    ex = sum(int(r[iex]) for r in body[b:b + 20]); s = sum(int(r[isamp]) for r in body[b:b + 20])
    if ex * 100.0 / tot_ex > 0.8 or s * 100.0 / max(tot_s, 1) > 0.8:
        print(f"{b:5d} {100.0 * ex / tot_ex:5.1f}% ex {100.0 * s / max(tot_s, 1):5.1f}% smp   {body[b][isrc].strip()[:60]}")
