#!/usr/bin/env python
"""cuobjdump -sass of libcvvae_b200.so -> per-kernel counts of the Blackwell mnemonics that prove the tcgen05 / TMEM /
TMA path (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops) plus the first occurrences as a listing.

    python tools/sass_summary.py > profiles/r02_sass_conv_tc.txt
"""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cvvae_b200", "lib", "libcvvae_b200.so")
MNEMONICS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "SYNCS", "UBLKCP", "HMMA", "FFMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    sys.path.insert(0, ROOT)
    import bench
    print(f"# cuobjdump -sass cvvae_b200/lib/libcvvae_b200.so   (sm_100a; source hash {bench.source_hash()}, "
          f"library sha256 {hashlib.sha256(open(LIB, 'rb').read()).hexdigest()[:16]})")
    print("# per kernel: instruction count of each mnemonic (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld TMEM->registers,")
    print("# UTMALDG/UTMASTG = TMA tensor load/store, UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier arrive/try_wait)")
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            kernels[cur]["_lines"] = []
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["_n"] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + "."):
                kernels[cur][mn] += 1
                if mn in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR") and len(kernels[cur]["_lines"]) < 12:
                    kernels[cur]["_lines"].append(line.strip()[:150])
        if ".2CTA" in op and op.startswith("UTCHMMA"):
            kernels[cur]["UTCHMMA.2CTA"] += 1
    try:
        names = subprocess.run(["cu++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    except Exception:
        names = list(kernels)
    for (k, c), nm in zip(kernels.items(), names):
        nm = nm.replace("cvvae::", "").replace("void ", "")
        cut = nm.find("(CUtensorMap")
        if cut < 0:
            cut = nm.find("(", nm.rfind(">") + 1) if ">" in nm else nm.find("(")
        nm = nm[:cut] if cut > 0 else nm
        row = "  ".join(f"{mn}={c[mn]}" for mn in MNEMONICS if c[mn])
        print(f"\n{nm}  [{c['_n']} instructions]\n    {row}")
        for ln in c["_lines"]:
            print("      " + ln)


if __name__ == "__main__":
    main()
