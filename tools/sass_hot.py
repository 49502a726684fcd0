"""Static look at a kernel's SASS: backward branches (loops) and the opcode mix of an address range.

    python tools/sass_hot.py <object.o> <mangled-kernel-substring> [lo hi]
"""
import collections
import re
import subprocess
import sys


def load(obj, sub):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    ins, cur, name = [], False, None
    for l in out.split("\n"):
        m = re.match(r"\s+Function : (\S+)", l)
        if m:
            cur = sub in m.group(1)
            name = m.group(1) if cur else name
            continue
        if cur:
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
    return name, ins


def hist(ins, lo, hi):
    c = collections.Counter()
    for a, t in ins:
        if lo <= a < hi:
            t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
            c[t2.split()[0].split(".")[0]] += 1
    return c


def main():
    obj, sub = sys.argv[1], sys.argv[2]
    name, ins = load(obj, sub)
    print(name, len(ins), "instructions")
    loops = []
    for a, t in ins:
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            loops.append((int(m.group(1), 16), a))
    for lo, hi in loops:
        print(f"  loop {lo:#x}..{hi:#x}: {(hi - lo) // 16 + 1} instructions")
    if len(sys.argv) >= 5:
        lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
        c = hist(ins, lo, hi + 16)
        tot = sum(c.values())
        fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
        print(f"range {lo:#x}..{hi:#x}: {tot} instructions, {fp64} FP64 ({100 * fp64 / tot:.1f} %)")
        for k, v in c.most_common(30):
            print(f"   {k:10s} {v:5d} {100 * v / tot:5.1f}%")


if __name__ == "__main__":
    main()
