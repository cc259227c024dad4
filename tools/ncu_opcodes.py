#!/usr/bin/env python
"""Aggregates the source page of an ncu report: executed warp instructions per opcode, shared-memory
wavefronts, stall samples.  usage: ncu_opcodes.py source.csv [n_batches]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
nb = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
ops = defaultdict(float)
wf = defaultdict(float)
wfi = defaultdict(float)
samples = defaultdict(float)
total = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stalls = defaultdict(float)
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]].strip()
    parts = src.split()
    if not parts:
        continue
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.split(".")[0]
    n = float(r[col["Instructions Executed"]] or 0)
    ops[op] += n
    total += n
    wf[op] += float(r[col["L1 Wavefronts Shared"]] or 0)
    wfi[op] += float(r[col["L1 Wavefronts Shared Ideal"]] or 0)
    samples[op] += float(r[col["# Samples"]] or 0)
    for s in stall_cols:
        stalls[s] += float(r[col[s]] or 0)
print(f"total executed warp instructions: {total:.0f}  per batch: {total / nb:.1f}")
ts = sum(samples.values())
for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{op:12s} {n / nb:10.1f}  {100 * n / total:5.1f}%   samples {100 * samples[op] / ts:5.1f}%   smem wavefronts {wf[op] / nb:8.1f} (ideal {wfi[op] / nb:8.1f})")
print("stall samples:")
tot = sum(stalls.values())
for s, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:12]:
    print(f"  {s:28s} {100 * v / tot:5.1f}%")
