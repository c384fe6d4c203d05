#!/usr/bin/env python
"""Regenerates profiles/r2_sass_evidence.txt: SASS excerpts of the built library that show which Blackwell
facilities the hot kernels use (bulk-copy engine + mbarrier in the SpMV, cluster barriers in the V-cycle tail,
cooperative grid barriers in the K-cycle and the sparse Cholesky), so the claims do not depend on a stray .so.

  python tools/sass_evidence.py        # needs cuobjdump (CUDA toolkit) and the built library
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sim3opt_b200", "lib", "libsim3opt_b200.so")
OUT = os.path.join(ROOT, "profiles", "r2_sass_evidence.txt")
KERNELS = {
    "spmv4_kernel": ["UBLKCP", "SYNCS", "LDS", "DFMA", "LDG", "STG"],
    "amg_tail_kernel": ["UCGABAR", "DFMA", "LDG"],
    "amg_coop_kernel": ["DFMA", "LDG", "ATOM", "RED", "MEMBAR"],
    "direct_kernel": ["DFMA", "MUFU", "BAR", "ATOM", "MEMBAR"],
    "linearize_kernel": ["DFMA", "DMUL", "DADD", "MUFU", "STL", "LDL"],
    "pcg_update_kernel": ["DFMA", "SHFL", "LDG", "STG"],
}


def main():
    if not os.path.exists(LIB):
        sys.exit(f"{LIB} missing: build it first (python -c 'import __graft_entry__ as g; g.build()')")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(line)
    lines = [f"SASS evidence for {os.path.relpath(LIB, ROOT)} (cuobjdump -sass), architectures: {', '.join(arch)}",
             f"{len(funcs)} device functions", ""]
    for key, mnems in KERNELS.items():
        for name, body in funcs.items():
            if key not in name:
                continue
            ops = collections.Counter()
            for ln in body:
                m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
                if m:
                    ops[m.group(1).split(".")[0]] += 1
            demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            lines.append(f"== {demangled[:150]}")
            lines.append("   instructions: %d   " % sum(ops.values()) + "  ".join(f"{k}={ops[k]}" for k in mnems if ops[k]))
            shown = 0
            for ln in body:
                if any(t in ln for t in ("UBLKCP", "SYNCS", "UCGABAR", "BAR.SYNC", "ERRBAR", "CCTL")) and shown < 6:
                    lines.append("   " + ln.strip()[:140])
                    shown += 1
            lines.append("")
    open(OUT, "w").write("\n".join(lines) + "\n")
    print("wrote", OUT, len(lines), "lines")


if __name__ == "__main__":
    main()
