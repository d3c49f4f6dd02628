"""Print the SASS of one kernel (substring match on the mangled name) from a cubin / executable, one instruction per
line, optionally only the innermost loop that contains a given opcode.  Development aid.
  python scripts/sass_fn.py <binary> <name-substring> [--loop OPCODE]"""
import re
import subprocess
import sys

binary, name = sys.argv[1], sys.argv[2]
loop_op = sys.argv[sys.argv.index("--loop") + 1] if "--loop" in sys.argv else None
out = subprocess.run(["cuobjdump", "-sass", binary], capture_output=True, text=True).stdout
cur, fn = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        fn[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m and cur:
        fn[cur].append((int(m.group(1), 16), m.group(2).strip()))
hits = [k for k in fn if name in k]
if not hits:
    sys.exit("no function matches; have:\n" + "\n".join(fn))
ins = fn[hits[0]]
print("#", hits[0], len(ins), "instructions")
if loop_op:
    # backward branches define loops [target, branch]; pick the smallest one containing the opcode
    best = None
    for a, t in ins:
        m = re.search(r"BRA\s+(?:\w+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) <= a:
            lo, hi = int(m.group(1), 16), a
            body = [x for x in ins if lo <= x[0] <= hi]
            if any(loop_op in x[1] for x in body) and (best is None or len(body) < len(best)):
                best = body
    ins = best or []
    print("# loop:", len(ins), "instructions")
    import collections
    c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in ins)
    print("#", dict(c.most_common()))
for a, t in ins:
    print(f"{a:05x}  {t}")
