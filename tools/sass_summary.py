"""Per-kernel SASS mnemonic histogram of libzkdl_b200.so (cuobjdump -sass), for profiles/: which pipes a kernel's instructions go to,
and the sm_100a-specific opcodes (UTCIMMA = tcgen05.mma, UTMALDG = TMA tile load, LDTM = tcgen05.ld, UTCBAR / SYNCS = mbarrier).
   python tools/sass_summary.py zkdl_b200/libzkdl_b200.so out.md [kernel-name-regex]"""
import collections
import re
import subprocess
import sys

so, out = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "").replace("zk::", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        op = m.group(1)
        base = op.split(".")[0]
        key = op if base in ("IMAD", "UTCIMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "IMMA", "LDG", "STG", "LDS", "STS") and len(op.split(".")) > 1 and base in ("IMAD", "UTMALDG", "LDG", "STG") else base
        if base == "IMAD":
            key = "IMAD.WIDE" if ".WIDE" in op else ("IMAD.HI" if ".HI" in op else ("IMAD.MOV/SHL" if (".MOV" in op or ".SHL" in op or ".IADD" in op) else "IMAD"))
        if base in ("LDG", "STG", "LDS", "STS"):
            key = base + (".128" if ".128" in op else (".64" if ".64" in op else ""))
        hist[kern][key] += 1
with open(out, "w") as f:
    f.write("# SASS mnemonic histogram per kernel (`cuobjdump -sass zkdl_b200/libzkdl_b200.so`, sm_100a)\n\n"
            "Static instruction counts (not executed counts).  IMAD.WIDE = 32x32->64 multiply-add (the limb product of the Montgomery\n"
            "chains), UTCIMMA = tcgen05.mma kind::i8, UTMALDG = TMA tensor tile load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,\n"
            "SYNCS = mbarrier, IMMA = mma.sync int8.\n\n| kernel | total | top opcodes |\n|---|---:|---|\n")
    for k, h in hist.items():
        if pat and not pat.search(k):
            continue
        tot = sum(h.values())
        top = ", ".join(f"{o} {c}" for o, c in h.most_common(9))
        special = ", ".join(f"**{o} {c}**" for o, c in h.items() if o.split(".")[0] in ("UTCIMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "IMMA", "UTCATOMSWS") )
        f.write(f"| `{k}` | {tot} | {top}{' — ' + special if special else ''} |\n")
print(open(out).read()[:2500])
