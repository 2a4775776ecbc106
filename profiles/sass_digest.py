"""Per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass): the evidence that the cubins are sm_100a, that the
hot loops use Blackwell's packed fp32 (FFMA2 / FMUL2 / FADD2), 128-bit streaming loads and evict-first stores, and that there
are no tensor-core instructions (the path has no contraction).  Usage: python profiles/sass_digest.py > profiles/r02_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "neuralnetworklibrary_b200", "libretina_sm100.so")
WATCH = ["FFMA2", "FMUL2", "FADD2", "FFMA", "MUFU", "LDG.E.128", "LDG.E.NA.128", "STG.E.EF.128", "STG.E.128", "LDG.E.U8", "LDG.E.S8",
         "UBLKCP", "SYNCS", "HMMA", "UTCHMMA", "UTCMMA", "LDTM", "ATOM", "RED", "BAR.SYNC", "CCTL.IVALL", "CCTL.E.PF2", "LDL", "STL", "DADD", "DMUL",
         "ACQBULK", "LDG.E.STRONG", "NANOSLEEP"]


def main():
    elf = subprocess.run(["cuobjdump", "-lelf", LIB], stdout=subprocess.PIPE, text=True).stdout
    print("library: %s (%.1f MB)" % (os.path.relpath(LIB, ROOT), os.path.getsize(LIB) / 1e6))
    print("cubins :", ", ".join(sorted(set(re.findall(r"sm_\w+", elf)))), "(%d)" % len(elf.strip().splitlines()))
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    kernels[cur][w] += 1
    total = collections.Counter()
    for c in kernels.values():
        total.update(c)
    print("kernels:", len(kernels), " SASS instructions:", total["_total"])
    print("whole library: " + "  ".join("%s %d" % (w, total[w]) for w in WATCH if total[w]))
    print("tensor-core / TMEM instructions (HMMA, UTC*MMA, LDTM): %d  -- none expected: no contraction on this path" %
          (total["HMMA"] + total["UTCHMMA"] + total["UTCMMA"] + total["LDTM"]))
    print()
    demangle = subprocess.run(["cu++filt"] + list(kernels), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    rows = []
    for (name, c), pretty in zip(kernels.items(), demangle):
        depth, cut = 0, len(pretty)
        for i, ch in enumerate(pretty):   # drop the parameter list: the first '(' outside the template brackets
            if ch == "<":
                depth += 1
            elif ch == ">":
                depth -= 1
            elif ch == "(" and depth == 0:
                cut = i
                break
        pretty = pretty[:cut].replace("void ", "").replace("(int)", "").replace("(bool)", "")
        rows.append((pretty[:78], c))
    print("%-78s %8s  %s" % ("kernel", "instr", "watched mnemonics"))
    for pretty, c in rows:
        print("%-78s %8d  %s" % (pretty, c["_total"], " ".join("%s:%d" % (w, c[w]) for w in WATCH if c[w])))


if __name__ == "__main__":
    main()
