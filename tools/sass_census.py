#!/usr/bin/env python
"""SASS census of the shipped library: per kernel, the instruction counts that show what it is made of —
UBLKCP (TMA 1-D bulk copies), SYNCS (mbarrier arrive / try_wait), fp64 arithmetic (DFMA / DADD / DMUL), vector
loads and stores, and the griddepcontrol instructions of programmatic dependent launch.
    python tools/sass_census.py > profiles/sass_census.txt      (needs cuobjdump; no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "domain-decomposed-pde-solver_b200", "lib", "libheat_b200.so")
WATCH = ["UBLKCP", "SYNCS", "DFMA", "DADD", "DMUL", "MUFU", "LDG", "STG", "LDS", "STS", "ATOM", "ATOMG", "RED", "SHFL", "BAR", "ACQBULK", "PREEXIT", "CCTL", "MEMBAR", "FENCE", "ERRBAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    kernels, cur = collections.OrderedDict(), None
    arch = set(re.findall(r"arch = (sm_\w+)", out))
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels, cubin arch {sorted(arch)}")
    tot = collections.Counter()
    for name, c in kernels.items():
        tot.update(c)
    print("# whole library: " + ", ".join(f"{k} {tot[k]}" for k in WATCH if tot[k]))
    print(f"# {'kernel':<110} {'insts':>6} " + " ".join(f"{w:>6}" for w in WATCH[:10]))
    for name, c in sorted(kernels.items(), key=lambda kv: -sum(kv[1].values())):
        dn = demangle(name)
        dn = re.sub(r"\(.*", "", dn)[:108]
        print(f"  {dn:<110} {sum(c.values()):>6} " + " ".join(f"{c[w]:>6}" for w in WATCH[:10]))


if __name__ == "__main__":
    sys.exit(main())
