"""Per-kernel count of the SASS instructions that prove the Blackwell data path (run here, no GPU needed):

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt

UTCHMMA = tcgen05.mma (".2CTA" = cta_group::2), LDTM / STTM = tcgen05.ld / .st (TMEM), UTMALDG / UTMASTG = TMA tensor
loads / stores, UTCBAR = tcgen05.commit, SYNCS = mbarrier traffic, HMMA = legacy mma.sync (none in the default build).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kddcup_2020_multimodalitiesrecall_2nd_place_b200", "libmmrecall.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "MUFU"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    names = re.findall(r"Function : (\S+)", sass)
    if names:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
        demangle = dict(zip(names, out))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\(.*", "", demangle.get(m.group(1), m.group(1)))
            cur = re.sub(r"^void ", "", cur).replace("mmr::", "")
            counts.setdefault(cur, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            base = op.split(".")[0]
            if base in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "MUFU"):
                counts[cur][base] += 1
                if base == "UTCHMMA" and ".2CTA" in op:
                    counts[cur]["UTCHMMA.2CTA"] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a), instruction counts per kernel; kernels without any "
          f"of these opcodes are omitted")
    print(f"{'kernel':<72}" + "".join(f"{o:>13}" for o in OPS))
    total = collections.Counter()
    for name, c in counts.items():
        if not any(c[o] for o in OPS if o not in ("MUFU", "SYNCS")):
            continue
        print(f"{name[:71]:<72}" + "".join(f"{c[o]:>13}" for o in OPS))
        total.update(c)
    print(f"{'TOTAL':<72}" + "".join(f"{total[o]:>13}" for o in OPS))


if __name__ == "__main__":
    sys.exit(main())
