"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use in the shipped library
(profiles/r02_sass_summary.md):   python tools/sass_summary.py [out.md]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "showtell_b200", "libshowtell_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "UTMALDG", "UBLKCP", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "UTCCP", "MULTIMEM", "REDUX"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_summary.md")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::|st::", "", cur)
            cur = re.sub(r"\(.*$", "", cur)[:90]
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    per[cur][k + (".2CTA" if ".2CTA" in op else "")] += 1
                    break
    cols = sorted({k for c in per.values() for k in c})
    tot = collections.Counter()
    lines = ["# SASS evidence: tcgen05 / TMEM / TMA / multimem instructions per kernel of libshowtell_b200.so", "",
             "`cuobjdump -sass showtell_b200/libshowtell_b200.so` (sm_100a), counted per kernel by `tools/sass_summary.py`. "
             "UTCHMMA = tcgen05.mma kind::f16 / tf32, UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk, "
             "LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops.", "",
             "| kernel | " + " | ".join(cols) + " |", "|---|" + "---:|" * len(cols)]
    for k, c in per.items():
        if not any(c[x] for x in cols if not x.startswith("SYNCS")):
            continue
        lines.append(f"| `{k}` | " + " | ".join(str(c[x]) if c[x] else "" for x in cols) + " |")
        tot.update(c)
    lines.append("| **total** | " + " | ".join(str(tot[x]) for x in cols) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[-12:]))


if __name__ == "__main__":
    main()
