"""Static SASS instruction counts per kernel of libg2048.so (evidence for tcgen05 / TMEM / bulk-copy use).
Usage: python tools/sass_evidence.py > profiles/r01_sass_evidence.txt"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "2048-ppo_b200/g2048/libg2048.so"
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "SYNCS", "ACQBULK", "FFMA", "VABSDIFF4", "VIMNMX", "PRMT", "LDS", "STG",
        "LOP3", "SHF", "IMAD", "POPC"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, order, cur = collections.defaultdict(collections.Counter), [], None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"^_ZN5g2048\d*(?:[a-z]{2}\d+)?", "", m.group(1))
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        counts[cur]["total"] += 1
        counts[cur][m.group(1)] += 1
print("# SASS evidence (cuobjdump -sass 2048-ppo_b200/g2048/libg2048.so, sm_100a), instruction counts per kernel:")
print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc,")
print("# UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops, VABSDIFF4 = 4-way byte abs-diff, VIMNMX = (packed) integer")
print("# min/max, PRMT = byte permute, ACQBULK = emitted once by each kernel that executes griddepcontrol.wait (programmatic dependent launch).\n# (static counts: a tcgen05.mma inside the k-step loop counts once)\n")
for k in order:
    c = counts[k]
    print(f"{k[:72]:72s} total={c['total']:5d}  " + " ".join(f"{n}={c[n]}" for n in KEYS if c[n]))
