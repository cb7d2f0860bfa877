"""SASS evidence for the hot kernels of libnbody_b200.so: opcode histogram per kernel and an excerpt of the densest
FP64 / packed-FP32 region (runs in the build container, no GPU).  python tools/sass_excerpt.py > profiles/r2_sass_excerpts.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nbodysimproject_b200", "libnbody_b200.so")
KERNELS = [
    ("ensemble_main_kernel<3, yoshida4> (thread per system; C3)", r"ensemble_main_kernelILi3ELi1ELb1ELb0", "DFMA"),
    ("ensemble_main_kernel<8, yoshida4> (C3)", r"ensemble_main_kernelILi8ELi1ELb1ELb0", "DFMA"),
    ("ensemble_main_kernel<4, whfast> (C4)", r"ensemble_main_kernelILi4ELi2ELb1ELb0", "DMUL"),
    ("hamsoft_run_kernel<3> (C1)", r"hamsoft_run_kernelILi3E", "DFMA"),
    ("largeN_accel_x2_kernel<IPT=2, MINB=4, no sums, eps>0> (C5 force, shipped default)", r"largeN_accel_x2_kernelILi2ELi4ELb0ELb0", "FFMA2"),
    ("largeN_pass_kernel<DENSITY> (C5 eps* sweeps)", r"largeN_pass_kernelILi0E", "MUFU.EX2"),
    ("mid_run_kernel<yoshida4> (9..64 bodies)", r"mid_run_kernelILi1ELb0E", "DFMA"),
]


def functions():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, body = None, {}
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            body[cur].append(ln)
    return body


def opcode(ln):
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    return m.group(1) if m else None


def main():
    body = functions()
    for title, rx, key in KERNELS:
        names = [n for n in body if re.search(rx, n)]
        if not names:
            print(f"== {title}: not found\n")
            continue
        name = names[0]
        lines = body[name]
        ops = collections.Counter(filter(None, (opcode(l) for l in lines)))
        fam = collections.Counter()
        for o, c in ops.items():
            fam[o.split(".")[0] if not o.startswith("MUFU") else o] += c
        print(f"== {title}\n   {name}\n   {len(lines)} SASS instructions; opcode families: "
              + ", ".join(f"{o} {c}" for o, c in fam.most_common(16)))
        notable = ["DFMA", "DMUL", "DADD", "MUFU.RSQ64H", "MUFU.RCP64H", "MUFU.RSQ", "MUFU.EX2", "FFMA2", "FMUL2", "FADD2", "UBLKCP",
                   "SYNCS", "SHFL", "LDS", "STS", "LDL", "STL", "CALL"]
        print("   notable: " + ", ".join(f"{o} {fam.get(o, 0)}" for o in notable if fam.get(o, 0)))
        # excerpt: the 40-instruction window with the most `key` opcodes
        hits = [1 if (opcode(l) or "").startswith(key) else 0 for l in lines]
        best, bi = -1, 0
        w = 40
        s = sum(hits[:w])
        for i in range(0, max(1, len(lines) - w)):
            if s > best:
                best, bi = s, i
            s += (hits[i + w] if i + w < len(lines) else 0) - hits[i]
        print(f"   densest {key} window ({best} of {w}):")
        for l in lines[bi:bi + w]:
            print("     " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip())
        print()


if __name__ == "__main__":
    main()
