#!/usr/bin/env python3
"""Instruction-footprint and stall summary from the source page of an ncu report (needs --import-source on / --set full):
which stall reasons the warp samples fall into, how many SASS instructions carry 90 % / 99 % of the executed instructions,
how they are spread over the kernel's code, and where the `no_instruction` (instruction-fetch) stalls sit.

usage: tools/ncu_source_footprint.py <file.ncu-rep> <kernel-name-regex> [...]
"""
import csv
import io
import subprocess
import sys

import numpy as np

rep = sys.argv[1]
for name in sys.argv[2:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + name], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not hdr:
        print(name, ": no source page")
        continue
    kname = next((r[1] for r in rows[:hdr[0]] if r and r[0] == "Kernel Name"), name)
    h = rows[hdr[0]]
    end = hdr[1] - 1 if len(hdr) > 1 else len(rows)
    data = [r for r in rows[hdr[0] + 1:end] if len(r) == len(h)]
    ci = {n: i for i, n in enumerate(h)}
    samples = sum(int(r[ci["# Samples"]]) for r in data)
    stalls = {k: sum(int(r[ci[k]]) for r in data) for k in h if k.startswith("stall_") and "Not Issued" not in k}
    e = np.array([int(r[ci["Instructions Executed"]]) for r in data])
    s = np.sort(e)[::-1]
    cs = np.cumsum(s) / max(1, s.sum())
    n90, n99 = int((cs < 0.9).sum()), int((cs < 0.99).sum())
    hot = np.where(e >= s[min(n99, len(s) - 1)])[0]
    ni = np.array([int(r[ci["stall_no_inst"]]) for r in data])
    print(f"== {kname[:100]} (first captured launch)")
    print(f"SASS instructions: {len(data)} ({len(data) * 16 / 1024:.0f} KB), never executed in this launch: {int((e == 0).sum())}")
    print(f"instructions carrying 90 % / 99 % of the executed instructions: {n90} / {n99} ({n99 * 16 / 1024:.0f} KB)")
    print(f"128-byte lines touched by the 99 % set: {len(set(hot // 8))} ({len(set(hot // 8)) * 128 / 1024:.1f} KB), spread over "
          f"{(hot.max() - hot.min()) * 16 / 1024:.0f} KB of code")
    print(f"warp-state samples: {samples}; share per stall reason:")
    for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:9]:
        print(f"   {k:28s} {100.0 * v / max(1, samples):5.1f} %")
    print("hot (99 % set) instructions per 8 KB of code:", np.bincount(hot // 512, minlength=(len(e) + 511) // 512).tolist())
    print("no_instruction samples per 8 KB of code:    ", np.add.reduceat(ni, np.arange(0, len(ni), 512)).tolist())
    print()
