"""profiles/r02_sass_extract.md: per-kernel counts of the SASS mnemonics that prove Blackwell-native code in libgradtts_b200.so
(tcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG, tcgen05.commit -> UTCBAR; /opt/skills/guides/B200_PROFILING.md)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "grad-tts_b200", "libgradtts_b200.so")
OUT = os.path.join(ROOT, "profiles", "r02_sass_extract.md")
PATS = collections.OrderedDict([
    ("UTCHMMA", r"\bUTCHMMA"), ("of which .2CTA", r"UTCHMMA\.2CTA"), ("UTCBAR", r"\bUTCBAR"), ("UTMALDG", r"\bUTMALDG"),
    ("of which .2CTA/.MULTICAST", r"UTMALDG\.[0-9]D\.(2CTA|MULTICAST)"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
    ("MUFU.EX2", r"MUFU\.EX2"), ("MUFU.RCP", r"MUFU\.RCP"), ("F{FMA,ADD,MUL}2", r"\bF(FMA|ADD|MUL)2\b"),
    ("STG.E.ENL2.256", r"STG\.E\.ENL2\.256"), ("HMMA (mma.sync)", r"\bHMMA\b")])

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for k, p in PATS.items():
            if re.search(p, line):
                counts[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()


def short(d):
    d = d.replace("(anonymous namespace)::", "").replace("gtts::", "").replace("void ", "")
    return re.sub(r"\(.*", "", d)


rows = collections.OrderedDict()
for fn, d in zip(counts, names):
    c = counts[fn]
    if not any(c[k] for k in ("UTCHMMA", "UTMALDG", "LDTM", "STTM")):
        continue
    key = re.sub(r"<.*", "", short(d))
    agg = rows.setdefault(key, [0, collections.Counter()])
    agg[0] += 1
    agg[1].update(c)
keys = list(PATS)
out = ["# SASS extract of grad-tts_b200/libgradtts_b200.so (round 2)", "",
       "`cuobjdump -sass grad-tts_b200/libgradtts_b200.so` (sm_100a only), counts of the mnemonics that prove Blackwell-native code, summed over the",
       "template instantiations of each tensor-core kernel.  tcgen05.mma -> `UTCHMMA` (`.2CTA` = cta_group::2), tcgen05.commit -> `UTCBAR`,",
       "TMA -> `UTMALDG`, tcgen05.ld/st -> `LDTM`/`STTM`; `HMMA` would be the legacy mma.sync path.  Regenerate with `python tools/sass_extract.py`.", "",
       "| kernel (instantiations) | " + " | ".join(keys) + " |", "|---|" + "---|" * len(keys)]
tot = collections.Counter()
for k, (n, c) in rows.items():
    out.append(f"| `{k}` ({n}) | " + " | ".join(str(c[x]) for x in keys) + " |")
    tot.update(c)
out.append("| **all tensor-core kernels** | " + " | ".join(str(tot[x]) for x in keys) + " |")
allk = collections.Counter()
for c in counts.values():
    allk.update(c)
out += ["", f"Whole library ({len(counts)} kernels): " + ", ".join(f"{k} {allk[k]}" for k in keys) + "."]
# the legacy mma.sync path, by kernel: only where a tcgen05 tile pipeline does not pay (K of 18 / 27, 32 x 32 head matrices) or as a
# selectable fallback next to the tcgen05 kernel
hm = collections.OrderedDict()
for fn, d in zip(counts, names):
    if counts[fn]["HMMA (mma.sync)"]:
        key = re.sub(r"<.*", "", short(d))
        a = hm.setdefault(key, [0, 0]); a[0] += 1; a[1] += counts[fn]["HMMA (mma.sync)"]
out += ["", "`HMMA` (mma.sync) by kernel: " + ", ".join(f"`{k}` ({n}): {c}" for k, (n, c) in hm.items()) + ".",
        "`first_conv_mma_kernel` is the default first conv of the bf16 mode (K = 18 | 27: 2-4 k-steps per pixel group, too thin for a tcgen05",
        "tile pipeline); the others are the mma.sync attention / weight-gradient kernels that the tcgen05 versions replaced on the default path",
        "(`attn_xk_tc_kernel`, `wgrad_tc_kernel`) and that remain for C = 256 contexts, stride-2 / transposed weight gradients and as fallbacks."]
open(OUT, "w").write("\n".join(out) + "\n")
sys.stdout.write("\n".join(out) + "\n")
