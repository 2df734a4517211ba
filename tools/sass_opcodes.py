#!/usr/bin/env python
"""SASS opcode summary of libfrg.so: which kernels hold tcgen05 / TMEM / TMA instructions (the mnemonics of
B200_PROFILING.md "What proves a Blackwell-native kernel").  Runs without a GPU.
    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "facerecognition_infrenceengine_b200", "csrc", "libfrg.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "UTCATOMSWS", "HMMA", "HGMMA",
         "SYNCS", "ELECT", "FMNMX", "FMNMX3", "REDG", "ATOMG", "ST.E.STRONG.SYS", "LD.E.STRONG.SYS", "ACQBULK", "NANOSLEEP"]


def main():
    elf = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    print("# %s" % os.path.relpath(LIB, ROOT))
    print("# embedded ELF images: %s" % ", ".join(sorted(set(re.findall(r"sm_\w+", elf)))))
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if cur and m:
            op = m.group(1)
            per[cur]["_n"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    per[cur][w] += 1
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print("# totals: " + ", ".join("%s=%d" % (w, total[w]) for w in WATCH if total[w]))
    print("%-34s %7s  %s" % ("kernel (demangled head)", "instrs", "watched opcodes"))
    names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    for (fn, c), nm in zip(per.items(), names):
        head = re.sub(r"^void frg::", "", nm)
        head = re.sub(r"\(.*", "", head)
        print("%-60s %7d  %s" % (head[:60], c["_n"], " ".join("%s=%d" % (w, c[w]) for w in WATCH if c[w])))


if __name__ == "__main__":
    sys.exit(main())
