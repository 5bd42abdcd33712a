"""Per-kernel counts of the Blackwell tensor-core / TMA / TMEM instructions in the shipped libnlc_b200.so (cuobjdump -sass):
the evidence that the contraction kernels are tcgen05 + TMEM + TMA (profiling recipe: B200_PROFILING.md, SASS mnemonics).
    python scripts/sass_extract.py > profiles/r02_sass_extract.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "diffusion-nlc_b200", "libnlc_b200.so")
# mnemonic prefix -> meaning
WHAT = collections.OrderedDict([
    ("UTCHMMA", "tcgen05.mma kind::f16 / kind::tf32 (UTCHMMA; .2CTA = cta_group::2)"),
    ("UTCQMMA", "tcgen05.mma kind::f8f6f4"),
    ("UTCBAR", "tcgen05.commit -> mbarrier"),
    ("LDTM", "tcgen05.ld (TMEM -> registers)"),
    ("STTM", "tcgen05.st (registers -> TMEM)"),
    ("UTCATOMSWS", "tcgen05.alloc / dealloc"),
    ("UTMALDG", "cp.async.bulk.tensor global -> shared (TMA load)"),
    ("UTMASTG", "cp.async.bulk.tensor shared -> global (TMA store)"),
    ("UTMAPF", "TMA descriptor prefetch"),
    ("SYNCS", "mbarrier arrive / try_wait"),
    ("HMMA", "legacy mma.sync (none expected)"),
])


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        for key in WHAT:
            if op.startswith(key):
                counts[cur][key] += 1
                if ".2CTA" in op:
                    counts[cur][key + ".2CTA"] += 1
    print("libnlc_b200.so (%d bytes), cuobjdump -sass, sm_100a; instruction counts per kernel (kernels with none of them "
          "omitted)\n" % os.path.getsize(LIB))
    for k, v in WHAT.items():
        print("  %-11s %s" % (k, v))
    print()
    total = collections.Counter()
    for fn in order:
        c = counts[fn]
        if not any(c[k] for k in WHAT if k != "SYNCS"):
            continue
        total.update(c)
        print(demangle(fn)[:150])
        print("    " + "  ".join("%s %d" % (k, c[k]) for k in sorted(c)))
    print("\nTOTAL  " + "  ".join("%s %d" % (k, total[k]) for k in sorted(total)))
    print("kernels in the library: %d" % len(order))


if __name__ == "__main__":
    sys.exit(main())
