"""Per-kernel counts of the SASS opcodes that prove a Blackwell-native path (B200_PROFILING.md): UTCHMMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / .st, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit, HMMA = mma.sync.
  python profiles/tools/sass_opcodes.py > profiles/r02_sass_opcodes.csv      (needs cuobjdump, no GPU)"""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "v-express_b200", "lib", "libvxb200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, cnt = None, collections.defaultdict(collections.Counter)
OPS = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "MUFU", "ACQBULK", "PREEXIT")
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if cur and m and m.group(1) in OPS:
        cnt[cur][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
print("kernel,UTCHMMA(tcgen05.mma),LDTM(tcgen05.ld),STTM(tcgen05.st),UTMALDG(TMA load),UTMASTG(TMA store),UTCBAR(tcgen05.commit),"
      "HMMA(mma.sync),MUFU,ACQBULK(griddepcontrol.wait),PREEXIT(griddepcontrol.launch_dependents)")
for nm, c in sorted((re.sub(r"\(.*", "", n).replace("vx::", "").replace("void ", ""), cnt[m]) for m, n in zip(cnt, names)):
    print(f'"{nm}",' + ",".join(str(c[o]) for o in OPS))
