#!/usr/bin/env python
"""Per-phase split of a kernel's warp-stall samples from an ncu report (source page): instructions are grouped into
contiguous segments by how often they execute (pair loop >> flush blocks >> culling >> per-unit code), each segment with
its share of samples, of executed warp instructions and its dominant stall reasons.
Usage: python tools/ncu_phases.py <file.ncu-rep> <kernel-regex> [top-n instructions]"""
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    # several kernels may match: take the first block
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    blk = rows[starts[0]:(starts[1] if len(starts) > 1 else len(rows))]
    print("kernel:", blk[0][1][:120])
    hdr = blk[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in blk[2:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
    seen, uniq = set(), []
    for r in data:          # the listing repeats itself when SASS is shown per source view
        if r[0] in seen:
            continue
        seen.add(r[0]); uniq.append(r)
    data = uniq

    def f(r, k):
        try:
            return int(r[ix[k]])
        except (ValueError, KeyError):
            return 0
    keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    tot = sum(f(r, "# Samples") for r in data) or 1
    tin = sum(f(r, "Instructions Executed") for r in data) or 1
    mx = max(f(r, "Instructions Executed") for r in data)
    segs, cur = [], None
    for i, r in enumerate(data):
        e = f(r, "Instructions Executed")
        cls = "hot" if e > 0.5 * mx else ("warm" if e > 0.12 * mx else ("mid" if e > 0.015 * mx else "rare"))
        if cur and cur["cls"] == cls:
            cur["hi"] = i; cur["s"] += f(r, "# Samples"); cur["e"] += e
            for k in keys:
                cur["st"][k] = cur["st"].get(k, 0) + f(r, k)
        else:
            cur = {"cls": cls, "lo": i, "hi": i, "s": f(r, "# Samples"), "e": e, "st": {k: f(r, k) for k in keys}}
            segs.append(cur)
    print(f"{len(data)} instructions, {tot} samples; classes by execution count relative to the hottest instruction ({mx})")
    for s in segs:
        if s["s"] / tot < 0.004:
            continue
        top = sorted(s["st"].items(), key=lambda kv: -kv[1])[:4]
        print(f'{s["cls"]:5s} [{s["lo"]:5d}-{s["hi"]:5d}] samples {100 * s["s"] / tot:5.1f}%  warp-instr {100 * s["e"] / tin:5.1f}%  '
              + ", ".join(f'{k.replace("stall_", "")} {100 * v / max(1, s["s"]):.0f}%' for k, v in top if v))
    if topn:
        for i in sorted(sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:topn]):
            r = data[i]
            st = {k.replace("stall_", ""): f(r, k) for k in keys if f(r, k) > 0.2 * f(r, "# Samples")}
            print(i, r[ix["Source"]].strip()[:70].ljust(70), f(r, "# Samples"), f(r, "Instructions Executed"), st)


if __name__ == "__main__":
    main()
