"""Show SASS context (with per-instruction samples and dominant stall reason) around the hottest instructions of an .ncu-rep."""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 4; ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 12
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ix = {k: i for i, k in enumerate(h)}; data = rows[2:]
stalls = [k for k in h if k.startswith("stall_")]
def dom(r):
    best = max(stalls, key=lambda k: float(r[ix[k]] or 0))
    return best if float(r[ix[best]] or 0) > 0 else ""
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:ntop]
for i in order:
    print("=" * 100)
    for j in range(max(0, i - ctx), min(len(data), i + ctx + 1)):
        r = data[j]
        print(("->" if j == i else "  "), r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(9), dom(r).ljust(18), r[ix["Source"]][:100])
