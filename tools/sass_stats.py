#!/usr/bin/env python
"""SASS evidence for the trace kernels (no GPU needed): disassembles one instantiation of
stb::trace_kernel from the in-tree library and prints (1) the instruction mix of the whole
function, (2) the innermost loop -- the backward branch with the most FP64 instructions in its
body -- with its per-opcode counts: the "instructions per ray-step" DESIGN.md quotes.

    python tools/sass_stats.py TRAJ EXTRAS HIST DIVC REFILL TILED CROSS STRAY LOOKUP BLOCK HOIST [--dump]
    python tools/sass_stats.py 1 0 0 1 0 1 0 0 0 256 1        # the bench kernel
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rayrtrace_b200", "_lib", "libsidtrace_b200.so")


def mangled(args):
    b = [f"Lb{int(a)}E" for a in args[:9]]
    return "_ZN3stb12trace_kernelI" + "".join(b) + f"Li{int(args[9])}ELb{int(args[10])}EEEvNS_7KParamsE"


def disasm(name):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", name, LIB], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def opcode(text):
    t = text.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0]


FP64 = ("DMUL", "DADD", "DFMA", "DSETP", "DMNMX")


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    # a mangled name may be given directly (e.g. _ZN3stb20endpoint_dual_kernelILb1ELb1EEEvNS_7KParamsE)
    name = args[0] if len(args) == 1 and args[0].startswith("_Z") else mangled(args)
    ins = disasm(name)
    if not ins:
        sys.exit(f"{name}: not found in {LIB}")
    print(f"# {name}\n# {len(ins)} SASS instructions in the function")
    # loops = backward branches; pick the body with the most FP64 instructions
    best = None
    for k, (addr, text) in enumerate(ins):
        m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?`\(\.L_x_\d+\)|BRA(?:\.\w+)*\s+(?:\w+,\s*)?0x([0-9a-f]+)", text)
        if opcode(text).startswith("BRA"):
            tm = re.search(r"0x([0-9a-f]+)", text)
            if not tm:
                continue
            tgt = int(tm.group(1), 16)
            if tgt < addr:
                body = [t for a, t in ins if tgt <= a <= addr]
                nfp = sum(opcode(t).split(".")[0] in FP64 for t in body)
                if best is None or nfp > best[0]:
                    best = (nfp, tgt, addr, body)
    whole = collections.Counter(opcode(t).split(".")[0] for _, t in ins)
    print("# whole function:", ", ".join(f"{k} {v}" for k, v in whole.most_common(12)))
    if best:
        nfp, tgt, addr, body = best
        c = collections.Counter(opcode(t) for t in body)
        print(f"# hottest loop 0x{tgt:x}..0x{addr:x}: {len(body)} instructions, {nfp} on the FP64 pipe")
        for k, v in sorted(c.items(), key=lambda kv: (-kv[1], kv[0])):
            print(f"{v:4d}  {k}")
        if "--dump" in sys.argv:
            for a, t in ins:
                if tgt <= a <= addr:
                    print(f"    /*{a:04x}*/  {t}")


if __name__ == "__main__":
    main()
