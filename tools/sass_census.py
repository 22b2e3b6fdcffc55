"""SASS census of libace_b200.so: per kernel, the counts of the Blackwell-native mnemonics (tcgen05.mma -> UTC*MMA, TMA ->
UTMALDG / UTMASTG, tcgen05.ld -> LDTM, tcgen05.commit / mbarrier -> UTCBAR / SYNCS, packed fp32 -> FFMA2) and of the legacy tensor
path (HMMA).  Run here (no GPU needed):  python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ace_b200", "lib", "libace_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "FFMA2", "FFMA", "HMMA", "MUFU.EX2", "STG", "LDG", "STS", "LDS", "ATOMG", "RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_n"] += 1
            for p in PAT:
                if op == p or op.startswith(p + "."):
                    kernels[cur][p] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    tot = collections.Counter()
    print(f"# SASS census of {os.path.relpath(LIB)} ({len(kernels)} kernels)")
    print("# columns: instructions | " + " ".join(PAT))
    for (name, c), dn in zip(kernels.items(), demangle):
        short = re.sub(r"ace::\(anonymous namespace\)::", "", dn)
        short = re.sub(r"\(ace::.*", "", short)[:150]
        print(f"{short}\n    {c['_n']:6d} | " + " ".join(f"{p}={c[p]}" for p in PAT if c[p]))
        tot.update(c)
    print("# TOTAL " + " ".join(f"{p}={tot[p]}" for p in PAT))


if __name__ == "__main__":
    sys.exit(main())
