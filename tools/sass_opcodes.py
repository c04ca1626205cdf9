"""Per-kernel SASS opcode counts of the shipped library (cuobjdump -sass), the evidence file profiles/rNN_sass_opcodes.txt.
    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "quantizations_b200", "libquantizations_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "LDTM", "STTM", "UTCBAR", "HMMA", "LDGSTS", "SYNCS", "ATOMS"]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return out if len(out) == len(names) else names


def short(n):
    n = n.replace("__nv_bfloat16", "bf16").replace("__half", "f16").replace("(bool)1", "true").replace("(bool)0", "false")
    n = re.sub(r"\(int\)(\d+)", r"\1", n)
    n = re.sub(r"^void ", "", n)
    cut = n.find(">(")
    return n[:cut + 1] if cut >= 0 else re.sub(r"\(.*$", "", n)


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {"instr": 0, **{o: 0 for o in OPS}}
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            kernels[cur]["instr"] += 1
            if op in kernels[cur]:
                kernels[cur][op] += 1
    names = list(kernels)
    pretty = [short(n) for n in demangle(names)]
    print("SASS opcode counts per kernel of quantizations_b200/libquantizations_b200.so (cuobjdump -sass, sm_100a) -- tools/sass_opcodes.py\n")
    print("UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG = cp.async.bulk.tensor (TMA tensor load), UBLKCP = cp.async.bulk, UBLKPF = cp.async.bulk.prefetch.L2,")
    print("UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, LDGSTS = cp.async, HMMA = mma.sync (legacy tensor path; the decode GEMVs use it on purpose: M = 1..8 tokens, the tensor")
    print("pipe only replaces FMA + shuffle work).  gemv_tc_kernel (tcgen05 GEMV) and its MT=16 batch instantiation are opt-in paths (DESIGN.md 4.1b).\n")
    w = max(len(p) for p in pretty) + 2
    print("kernel".ljust(w) + "".join(c.rjust(8) for c in ["instr"] + OPS))
    tot = {c: 0 for c in ["instr"] + OPS}
    for n, p in sorted(zip(names, pretty), key=lambda t: t[1]):
        k = kernels[n]
        print(p.ljust(w) + "".join(str(k[c]).rjust(8) for c in ["instr"] + OPS))
        for c in tot:
            tot[c] += k[c]
    print("TOTAL".ljust(w) + "".join(str(tot[c]).rjust(8) for c in ["instr"] + OPS))


if __name__ == "__main__":
    sys.exit(main())
