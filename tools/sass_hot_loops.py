"""cuobjdump -sass of the hot kernels -> profiles/sass_hot_loops_r02.txt: per kernel the instruction mix and the hot loop
(the inner loop holding the most POPCs), first 64 instructions. Runs without a GPU.

    python tools/sass_hot_loops.py > profiles/sass_hot_loops_r02.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "cbird_b200", "build")
# (object, kernel name part, longest loop body considered: keeps the pick on the inner loop, not on a loop around it)
WANT = [("mih.o", "mih_bucket_kernelILi1E", 800), ("mih.o", "mih2_bucket_kernel", 200), ("mih.o", "mih2_scatter_all_kernel", 200),
        ("dct_index.o", "find_small_kernelILi2E", 900), ("scan64.o", "scan64_kernelILi0E", 2000), ("scan64.o", "scan64_kernelILi2E", 2000)]
INS = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);")


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = INS.search(line)
        if m and cur:
            body[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return body


def opcode(text):
    t = text.split()
    op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
    return op.split(".")[0]


print("SASS of the round-2 hot loops (final build): cuobjdump -sass of cbird_b200/build/{mih,dct_index,scan64}.o (sm_100a), "
      "loop = the inner loop with the most POPCs (tools/sass_hot_loops.py)\n")
cache = {}
for obj, key, longest in WANT:
    if obj not in cache:
        cache[obj] = functions(obj)
    for name, ins in cache[obj].items():
        if key not in name:
            continue
        mix = collections.Counter(opcode(t) for _, t in ins)
        print("== " + name)
        print("   whole kernel: %d instructions; POPC %d, LOP3 %d, LDS %d, LDG %d, STG %d, ATOMS %d, ATOMG %d; UTMALDG %d, UTCMMA %d "
              "(no TMA / tcgen05 on this path, by design: DESIGN.md section 3)"
              % (len(ins), mix["POPC"], mix["LOP3"], mix["LDS"], mix["LDG"], mix["STG"], mix["ATOMS"], mix["ATOMG"] + mix["ATOM"],
                 mix["UTMALDG"], mix["UTCMMA"]))
        index = {a: i for i, (a, _) in enumerate(ins)}
        loops = []  # (first index, last index) of every backward branch
        for i, (a, t) in enumerate(ins):
            m = re.search(r"BRA\S*\s+(?:\S+,\s*)?(?:`\()?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a and int(m.group(1), 16) in index:
                loops.append((index[int(m.group(1), 16)], i))
        best = None
        for lo, hi in loops:  # the most POPCs among the loops of at most `longest` instructions, the shortest such loop
            body = ins[lo:hi + 1]
            if len(body) > longest:
                continue
            pop = sum(1 for _, x in body if opcode(x) == "POPC")
            if best is None or pop > best[0] or (pop == best[0] and len(body) < len(best[1])):
                best = (pop, body)
        if best and best[0]:
            body = best[1]
            bm = collections.Counter(opcode(t) for _, t in body)
            print("   hot loop 0x%x..0x%x (%d instructions): %s" % (body[0][0], body[-1][0], len(body), dict(bm)))
            for a, t in body[:64]:
                print("      /*%04x*/ %s" % (a, t))
            if len(body) > 64:
                print("      ... (%d more)" % (len(body) - 64))
        print()
