#!/usr/bin/env python
"""Counts the Blackwell-specific SASS mnemonics per kernel of the built library (runs on the CPU box):
    python tools/sass_counts.py > profiles/rNN_sass_counts.txt
UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store,
UBLKCP = cp.async.bulk, LDGSTS = cp.async, SYNCS = mbarrier ops, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "crb-active-3ddet_b200", "lib", "libcrb3d_sm100.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "UTCBAR", "UTCATOMSWS", "HMMA", "FFMA", "DFMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    counts = collections.OrderedDict()
    cur = None
    it = iter(names)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\((int|bool)\)", "", next(it)).replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
            cur = re.sub(r"\(.*", "", cur)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        if base in KEYS:
            counts[cur][base] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            counts[cur]["UTCHMMA.2CTA"] += 1
    total = collections.Counter()
    print("# SASS mnemonic counts per kernel of lib/libcrb3d_sm100.so (cuobjdump -sass, sm_100a); kernels without any of the")
    print("# tensor-core / TMA / async-copy mnemonics are summed in the last line. %d kernels in the library." % len(counts))
    print("%-64s " % "kernel" + " ".join("%12s" % k for k in KEYS[:10]))
    plain = 0
    for k, c in counts.items():
        total.update(c)
        if not any(c[x] for x in KEYS[:10]):
            plain += 1
            continue
        print("%-64s " % k[:64] + " ".join("%12d" % c[x] for x in KEYS[:10]))
    print("%-64s " % "TOTAL (all kernels)" + " ".join("%12d" % total[x] for x in KEYS[:10]))
    print("# %d kernels use none of them (SIMT integer / fp32 / fp64 kernels); FFMA total %d, DFMA total %d, HMMA (mma.sync) total %d" %
          (plain, total["FFMA"], total["DFMA"], total["HMMA"]))


if __name__ == "__main__":
    main()
