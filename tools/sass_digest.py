"""Per-kernel SASS digest of libmosaic_b200.so: counts of the mnemonics that prove the Blackwell paths
(UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG/UTMASTG = 2-D TMA, UBLKCP = bulk copy, SYNCS = mbarrier, IDP = dp2a/dp4a,
POPC, DADD/DMUL = the fp64 coordinate pipeline).  usage: python tools/sass_digest.py > profiles/rNN_sass_digest.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "IDP", "POPC", "DADD", "DMUL", "FFMA2", "VIMNMX", "BAR.SYNC", "STS", "LDS", "LDG", "STG"]


def main():
    so = ROOT / "mosaic-library_b200" / "libmosaic_b200.so"
    txt = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
    cur, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            total[cur] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    counts[cur][k] += 1
    print(f"# SASS digest of {so.name} (cuobjdump -sass, sm_100a); columns: instructions, then non-zero mnemonic counts")
    for fn, c in counts.items():
        print(f"{fn}: {total[fn]} instr; " + ", ".join(f"{k}={v}" for k, v in c.items() if v))


if __name__ == "__main__":
    main()
