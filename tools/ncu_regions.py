"""Bucket stall samples of an .ncu-rep by SASS region delimited by landmark opcodes; per-region dominant stall reasons."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; step = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ix = {k: i for i, k in enumerate(h)}; data = rows[2:]
stalls = [k for k in h if k.startswith("stall_")]
marks = ("LDGSTS", "STS.128", "UTCHMMA", "UTCBAR", "STG", "LDG", "UBLKCP", "LDTM", "BAR.SYNC", "SYNCS", "EXIT", "TANH", "LDS")
for b in range(0, len(data), step):
    blk = data[b:b + step]
    n = sum(int(r[ix["# Samples"]]) for r in blk)
    ex = max(int(r[ix["Instructions Executed"]] or 0) for r in blk)
    agg = collections.Counter()
    for r in blk:
        for k in stalls:
            agg[k] += float(r[ix[k]] or 0)
    ops = collections.Counter()
    for r in blk:
        for m in marks:
            if m in r[ix["Source"]]:
                ops[m] += 1
    top = ", ".join(f"{k[6:]}={int(v)}" for k, v in agg.most_common(3) if v > 0)
    print(f"{b:5d} samples={n:6d} maxexec={ex:9d}  [{top}]  {dict(ops)}")
