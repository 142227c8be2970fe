"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA load / store), HMMA (mma.sync), STAS (st.async, distributed shared memory),
F2FP.SATFINITE (the fp16 state's saturating pack), from `cuobjdump -sass` of the shipped .so.
  python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "rgqa_b200", "_C", "librgqa_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pat = {"UTC*MMA": r"\bUTC[A-Z]*MMA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UTMALDG": r"\bUTMALDG", "UTMASTG": r"\bUTMASTG",
       "HMMA": r"\bHMMA", "UTCBAR": r"\bUTCBAR", "SYNCS": r"\bSYNCS", "LDSM": r"\bLDSM", "FFMA2": r"\bFFMA2",
       "STAS": r"\bSTAS", "F2FP.SATFINITE": r"\bF2FP\.SATFINITE"}   # st.async into the peer CTA's shared memory; saturating fp16 pack
cur = None
counts = collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k, p in pat.items():
        if re.search(p, line):
            counts[cur][k] += 1
dem = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)}  (sm_100a) — instruction counts per kernel")
print("# " + " | ".join(["kernel"] + list(pat)))
tot = collections.Counter()
for (mangled, c), name in zip(counts.items(), dem):
    if not any(c[k] for k in ("UTC*MMA", "LDTM", "UTMALDG", "UTMASTG", "HMMA")):
        continue
    name = name.replace("rgqa::(anonymous namespace)::", "").replace("void ", "")
    name = re.sub(r"\(CUtensorMap.*$|\((?:int|float|double|long|const|__nv|unsigned|void|rgqa)[^<]*$", "", name)
    print(f"{name[:90]:90s} " + " ".join(f"{k}={c[k]}" for k in pat if c[k]))
    tot.update(c)
print("TOTAL " + " ".join(f"{k}={tot[k]}" for k in pat))
