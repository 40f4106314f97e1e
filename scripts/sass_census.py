"""Static proof that the kernels are Blackwell-native (no GPU needed): per kernel of libppasr_b200.so, how many SASS
instructions of each telling class `cuobjdump -sass` shows (B200_PROFILING.md "What proves a Blackwell-native kernel"):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, SYNCS = mbarrier, HMMA = legacy mma.sync (must
be 0).  python scripts/sass_census.py > profiles/<round>_sass_census.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ppasr_b200", "lib", "libppasr_b200.so")
CLASSES = [("UTC*MMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
           ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA"),
           ("FFMA2", r"\bFFMA2"), ("MUFU", r"\bMUFU")]

if __name__ == "__main__":
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    rows = []
    for block in re.split(r"\n\s*Function : ", sass)[1:]:
        name = block.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(.*", "", dem).replace("void ", "")
        cnt = collections.OrderedDict((k, len(re.findall(rx, block))) for k, rx in CLASSES)
        n = len(re.findall(r"^\s+/\*[0-9a-f]{4}\*/", block, flags=re.M))
        rows.append((dem[:72], n, cnt))
    rows.sort()
    print("# cuobjdump -sass ppasr_b200/lib/libppasr_b200.so  (sm_100a); instruction counts per kernel")
    print(f"{'kernel':72} {'instr':>6} " + " ".join(f"{k:>7}" for k, _ in CLASSES))
    for dem, n, cnt in rows:
        print(f"{dem:72} {n:>6} " + " ".join(f"{v:>7}" for v in cnt.values()))
    tot = collections.Counter()
    for _, _, cnt in rows:
        tot.update(cnt)
    print(f"{'TOTAL (' + str(len(rows)) + ' kernels)':72} {sum(r[1] for r in rows):>6} " + " ".join(f"{tot[k]:>7}" for k, _ in CLASSES))
