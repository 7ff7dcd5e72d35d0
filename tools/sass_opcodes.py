"""Per-kernel counts of the Blackwell-only SASS opcodes in libdfd.so (cuobjdump -sass): python tools/sass_opcodes.py > profiles/sass_opcodes_rNN.txt"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "real-time-video-deepfake-detection_b200", "libdfd.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
OPS = ("UTCHMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "REDUX", "FFMA2", "MUFU.EX2", "MUFU.TANH")
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for op in OPS:
            if re.search(r"\b" + re.escape(op) + r"\b", line):
                counts[cur][op] += 1
print("arch sm_100a; opcode counts per kernel (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA tensor load/store, "
      "UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit)")
print(f"{'kernel':100s} " + " ".join(f"{o:>9s}" for o in OPS))
for k, c in counts.items():
    if sum(c.values()) == 0:
        continue
    print(f"{k[:100]:100s} " + " ".join(f"{c[o]:9d}" for o in OPS))
