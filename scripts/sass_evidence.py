"""Regenerates profiles/r02_sass_tcgen05.txt: per-kernel SASS mnemonic counts of the tcgen05 engine (run after build())."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "open_l2o_b200", "csrc", "libl2o_b200.so")],
                     capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "MUFU.EX2", "MUFU.RCP", "FFMA", "STS", "LDS",
        "LDG", "STG", "BAR.SYNC", "ELECT", "USETMAXREG"]
out = ["SASS evidence of the tcgen05 / TMA path (cuobjdump -sass open_l2o_b200/csrc/libl2o_b200.so, sm_100a).",
       "Mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st,",
       "UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops, MUFU = activation pipe.", ""]
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    if not re.search(r"l2o(2tc|3tcb|4tcb2)", name) or "prep_weights" in name:
        continue
    lines = f.split("\n")
    cnt = collections.Counter()
    for line in lines:
        for k in keys:
            if re.search(r"\b" + re.escape(k), line):
                cnt[k] += 1
    regs = sorted({int(m) for m in re.findall(r"\bR(\d+)\b", f)})
    ninstr = sum(1 for line in lines if re.search(r"/\*[0-9a-f]{4,6}\*/", line))
    out.append("== %s\n   instructions: %d, highest register R%d\n   %s" % (
        name, ninstr, regs[-1] if regs else -1, "  ".join("%s=%d" % (k, cnt[k]) for k in keys if cnt[k])))
    ex = [line.strip() for line in lines if "UTCHMMA" in line][:4]
    out.append("   first tensor-core issues:\n" + "\n".join("      " + e[:120] for e in ex) + "\n")
open(os.path.join(ROOT, "profiles", "r02_sass_tcgen05.txt"), "w").write("\n".join(out))
print("wrote profiles/r02_sass_tcgen05.txt")
