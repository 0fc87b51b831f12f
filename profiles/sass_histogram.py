"""Instruction histogram of libfir_b200.so from `cuobjdump -sass` (run on the CPU box; no GPU needed).

usage: python profiles/sass_histogram.py [path/to/libfir_b200.so] > profiles/r2_sass_histogram.txt

Per kernel: the count of every opcode that proves the Blackwell data path (tcgen05 MMA = UTCHMMA, TMA loads = UTMALDG,
TMEM loads = LDTM, tcgen05.commit = UTCBAR, mbarrier = SYNCS, cp.async = LDGSTS, cluster barrier = UCGABAR) plus the
arithmetic mix (FFMA/FADD/FMUL/MUFU/DFMA/DADD/DMUL) of the CUDA-core kernels."""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "fast-image-recognition_b200", "libfir_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "UCGABAR", "LDGSTS", "LDG", "STG", "LDS", "STS",
       "ATOM", "ATOMS", "RED", "SHFL", "FFMA", "FFMA2", "FADD", "FMUL", "MUFU", "DFMA", "DADD", "DMUL", "HMMA", "IMAD", "BAR", "BRA"]
kern, hist, order = None, collections.defaultdict(collections.Counter), []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        order.append(kern)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
        if m.group(1) in ("UTCHMMA", "UTMALDG", "UTCBAR", "LDTM", "MUFU"):
            hist[kern][m.group(1) + m.group(2)] += 1


def demangle(n):
    try:
        return subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n


print("# cuobjdump -sass %s — per-kernel opcode counts (static instructions)" % os.path.basename(so))
print("# arch:", ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", out)))))
for k in order:
    h = hist[k]
    tot = sum(v for kk, v in h.items() if "." not in kk)
    name = demangle(k)
    print("\n== %s\n   total %d" % (name[:200], tot))
    print("   " + "  ".join("%s=%d" % (kk, h[kk]) for kk in KEY if h.get(kk)))
    detail = sorted((kk, v) for kk, v in h.items() if "." in kk)
    if detail:
        print("   " + "  ".join("%s=%d" % kv for kv in detail))
