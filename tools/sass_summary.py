"""SASS evidence (no GPU needed): per-kernel histogram of the Blackwell-specific opcodes in libregen_sm100.so.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt

cuobjdump -sass of the in-tree library; for every kernel: registers / shared memory (cuobjdump --dump-resource-usage) and
the counts of the opcodes that prove tcgen05 / TMEM / TMA use (mnemonics per /opt/skills/guides/B200_PROFILING.md):
  UTCHMMA / UTCQMMA  tcgen05.mma (kind::f16 / f8f6f4)        LDTM / STTM  tcgen05.ld / st (TMEM)
  UTMALDG / UTMASTG  TMA tensor load / store                  UTMAPF       TMA L2 prefetch
  UTCBAR             tcgen05.commit -> mbarrier               SYNCS        mbarrier try_wait / arrive
  UTCATOMSWS / UTCCP TMEM allocation management
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "regennet_b200", "csrc", "libregen_sm100.so")
CUOBJDUMP = "/usr/local/cuda/bin/cuobjdump"
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "FFMA",
        "LDG", "STG", "LDS", "STS", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run([CUOBJDUMP, "--dump-resource-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = " ".join(re.findall(r"(?:REG|SHARED|LOCAL|STACK):\d+", line))
            cur = None
    hist = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            hist[cur]["_total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    hist[cur][k] += 1
            if ".2CTA" in op:
                hist[cur]["*.2CTA"] += 1
            if "MULTICAST" in op:
                hist[cur]["*.MULTICAST"] += 1
    names = demangle(list(hist))
    total = collections.Counter()
    print("SASS summary of %s (sm_100a), one block per kernel; opcode counts are static instruction counts" % os.path.relpath(LIB, ROOT))
    print(subprocess.run([CUOBJDUMP, "--list-elf", LIB], capture_output=True, text=True).stdout.strip())
    print()
    for fn, c in hist.items():
        name = names.get(fn, fn)
        name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)", "{anon}"))
        cols = " ".join("%s=%d" % (k, c[k]) for k in KEYS + ["*.2CTA", "*.MULTICAST"] if c[k])
        print("%s\n    %s | %d instructions | %s" % (name, usage.get(fn, "?"), c["_total"], cols))
        total.update(c)
    print()
    print("library total: " + " ".join("%s=%d" % (k, total[k]) for k in KEYS + ["*.2CTA", "*.MULTICAST"] if total[k]))


if __name__ == "__main__":
    sys.exit(main())
