#!/usr/bin/env python
"""Reads a .ncu-rep (no GPU needed) and prints, for the profiled kernel, the SASS instructions executed
per warp-step and where the stall samples are: every instruction with its executed count (relative to
the hottest instruction = once per warp-step) and stall samples.

    python tools/ncu_hot.py gpurun_out/r02/prof_cross_static_divc_hoist.ncu-rep [--min-frac 0.2] [--all]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    min_frac = 0.2
    if "--min-frac" in sys.argv:
        min_frac = float(sys.argv[sys.argv.index("--min-frac") + 1])
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print("#", rows[0][1][:200])
    h = rows[1]
    ia, isrc, isamp, iexec = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    data = [r for r in rows[2:] if len(r) == len(h)]
    ex = [int(r[iexec]) for r in data]
    # "once per step" = the most common large count among FP64 instructions
    ref = max(ex)
    tot_inst = sum(ex)
    tot_samp = sum(int(r[isamp]) for r in data)
    print(f"# {len(data)} SASS instructions, {tot_inst} warp-instructions executed, hottest {ref} -> "
          f"{tot_inst / ref:.1f} instructions per hottest-instruction execution; {tot_samp} stall samples")
    acc = {}
    for r in data:
        for i, n in stall_cols:
            acc[n] = acc.get(n, 0) + int(r[i] or 0)
    print("# stall samples by reason:", ", ".join(f"{n[6:]} {v}" for n, v in sorted(acc.items(), key=lambda kv: -kv[1]) if v > tot_samp * 0.01))
    for r, e in zip(data, ex):
        if "--all" in sys.argv or e >= min_frac * ref:
            st = {n[6:]: int(r[i] or 0) for i, n in stall_cols}
            top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
            tops = " ".join(f"{k}:{v}" for k, v in top if v)
            print(f"{e / ref:6.3f} {int(r[isamp]):7d}  {r[isrc].strip():70s} {tops}")


if __name__ == "__main__":
    main()
