"""Static SASS instruction mix of the hot kernels in libgrid_b200.so (cuobjdump -sass): opcode
histogram per kernel, so that a reader can check what the kernels are made of (predicated DFMA,
LOP3 mask tests, shared-memory traffic, bulk-async / mbarrier instructions of the CTA-tile family)
without a GPU.  Usage: python tools/sass_mix.py [substring of the mangled kernel name ...]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "cp2k_b200", "csrc", "libgrid_b200.so")
DEFAULT = ["tiled_kernelILb1ELi0ELi2", "tiled_kernelILb0ELi0ELi2", "ctile_kernelILb1ELi0ELi2", "generic_kernelILb0",
           "coef_to_hab_kernelILi16", "pab_to_coef_kernelILi16", "peer_sum_kernel"]


def main():
    want = sys.argv[1:] or DEFAULT
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    name, mix, pred = None, {}, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(2)
            mix.setdefault(name, collections.Counter())[op.split(".")[0]] += 1
            if m.group(1) and op.startswith("DFMA"):
                pred.setdefault(name, [0])[0] += 1
    for w in want:
        for k in sorted(mix):
            if w in k:
                c = mix[k]
                total = sum(c.values())
                top = ", ".join(f"{op} {n}" for op, n in c.most_common(14))
                print(f"{k}\n  {total} instructions; predicated DFMA {pred.get(k, [0])[0]} of {c['DFMA']} DFMA\n  {top}")
                special = {op: n for op, n in c.items() if op in ("UBLKCP", "SYNCS", "USETMAXREG", "RED", "ATOM", "ATOMS",
                                                                  "SHFL", "LDS", "STS", "LDG", "STG", "R2P", "MUFU")}
                print("  " + ", ".join(f"{op} {n}" for op, n in sorted(special.items())))


if __name__ == "__main__":
    main()
