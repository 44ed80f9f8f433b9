"""Static view of a kernel's SASS: dumps it and summarises the outermost hot loop (largest backward branch span that
   contains a SYNCS.PHASECHK, i.e. the day loop).  python tools/sass_loop.py <obj> <mangled-name-substring> [out.sass]"""
import collections, re, subprocess, sys
obj, key = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, funcs = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
names = [n for n in funcs if key in n]
assert names, "no kernel matches"
ins = funcs[names[0]]
print(names[0], len(ins), "instructions")
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write("\n".join("%04x %s" % i for i in ins))
# backward branches
loops = []
for a, t in ins:
    m = re.search(r"BRA(?:\.U)?(?:\.ANY)? (?:!?U?P\d, )?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a:
            loops.append((tgt, a))
def has_wait(lo, hi):
    return any(lo <= a <= hi and "PHASECHK" in t for a, t in ins)
cands = [l for l in loops if has_wait(*l)]
# the day loop: smallest span containing the phase check and an STG
def has(lo, hi, s):
    return any(lo <= a <= hi and s in t for a, t in ins)
cands = sorted([l for l in cands if has(l[0], l[1], "STG")], key=lambda l: l[1] - l[0])
lo, hi = cands[0]
body = [(a, t) for a, t in ins if lo <= a <= hi]
print("day loop 0x%x-0x%x: %d instructions (static)" % (lo, hi, len(body)))
c = collections.Counter()
for a, t in body:
    op = t.split()[0]
    if op.startswith("@"):
        op = t.split()[1]
    c[op.split(".")[0]] += 1
print(", ".join("%s=%d" % kv for kv in c.most_common(40)))
