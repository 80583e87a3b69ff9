"""SASS mnemonic histogram per kernel of the built library (no GPU needed):

    python profiles/sass_histogram.py > profiles/r2_sass_histogram.md
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "selftoktokenizer_b200", "csrc", "libselftok_b200.so")
COLS = ["UTCHMMA", "UTMALDG", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "SYNCS", "HMMA", "MUFU.EX2", "F2FP", "ERRBAR", "MEMBAR",
        "NANOSLEEP", "FFMA2", "FFMA", "LDS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = {}
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_all"] += 1
            for c in COLS:
                if op == c or op.startswith(c + ".") or (c == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                    counts[cur][c] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    for k, d in zip(counts, dem):
        d = re.sub(r"^void ", "", d).replace("(anonymous namespace)::", "").replace("stk::", "")
        d = re.sub(r"\((\w+)\)(\d+)", r"\2", d)                    # <(bool)1, (int)3> -> <1, 3>
        d = re.sub(r"\(.*$", "", d)
        names[k] = d
    print("# SASS mnemonic histogram of `libselftok_b200.so` (final round-2 build, `cuobjdump -sass`, sm_100a)\n")
    print("Static instruction counts per kernel of the mnemonics that identify the Blackwell paths: `UTCHMMA` = tcgen05.mma kind::f16,\n"
          "`UTMALDG` = TMA tensor load, `LDTM` / `STTM` = tcgen05.ld / st, `UTCBAR` = tcgen05.commit, `UTCATOMSWS` = TMEM alloc / dealloc,\n"
          "`SYNCS` = mbarrier operations, `NANOSLEEP` = suspended mbarrier waits, `FFMA2` = packed fp32 FMA (fma.rn.f32x2).\n"
          "**No `HMMA` (mma.sync) anywhere in the library** (asserted by `tests/test_host_and_abi.py`).  Regenerate:\n"
          "`python profiles/sass_histogram.py > profiles/r2_sass_histogram.md` (needs no GPU).\n")
    print("| kernel | SASS instr | " + " | ".join(COLS) + " |")
    print("|---|---:|" + "---:|" * len(COLS))
    for k in sorted(counts, key=lambda k: -counts[k]["_all"]):
        c = counts[k]
        print(f"| `{names[k]}` | {c['_all']} | " + " | ".join(str(c[x]) if c[x] else "" for x in COLS) + " |")


if __name__ == "__main__":
    main()
