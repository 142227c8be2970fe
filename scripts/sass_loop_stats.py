"""Static instruction mix of a GEMM kernel's epilogue loop, from the built library (no GPU needed):

    python scripts/sass_loop_stats.py 'pair_kernelILi4E13__nv_bfloat16Lb1ELi8ELi3E'     # substring of the mangled name

Finds the smallest backward-branch loop that contains the first TMEM load (LDTM) and prints its length and opcode
histogram — the number the producer-epilogue clean-up of round 1 was steered by (820 -> 685 static instructions per
32-column slab iteration).  Registers / stack come from `cuobjdump -res-usage`."""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rgqa_b200", "_C", "librgqa_b200.so")


def main():
    pat = sys.argv[1] if len(sys.argv) > 1 else "pair_kernelILi4E13__nv_bfloat16Lb1ELi8ELi3E"
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    names = sorted({m for m in re.findall(r"Function (\S+?):", res) if pat in m})
    if not names:
        sys.exit(f"no kernel matching {pat!r} in {LIB}")
    for name in names:
        usage = re.search(re.escape(name) + r":\s*\n\s*(.*)", res)
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", name, LIB], capture_output=True, text=True).stdout
        lines = [l for l in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]
        addr = lambda l: int(re.match(r"\s+/\*([0-9a-f]{4,5})\*/", l).group(1), 16)
        ld = [addr(l) for l in lines if "LDTM" in l]
        print(name[:150])
        print("  ", usage.group(1).strip() if usage else "")
        if not ld:
            print("   no LDTM in this kernel")
            continue
        loops = []
        for l in lines:
            m = re.search(r"BRA 0x([0-9a-f]+)", l)
            if m and int(m.group(1), 16) < ld[0] < addr(l):
                loops.append((int(m.group(1), 16), addr(l)))
        lo, hi = min(loops, key=lambda x: x[1] - x[0])
        ops = collections.Counter()
        for l in lines:
            if lo <= addr(l) <= hi:
                ops[re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_]+)", l).group(2)] += 1
        print(f"   loop {lo:#x}..{hi:#x}: {sum(ops.values())} instructions (static)")
        print("   " + ", ".join(f"{k} {v}" for k, v in ops.most_common(16)))


if __name__ == "__main__":
    main()
