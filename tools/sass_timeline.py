"""Compressed timeline of a kernel's SASS: runs of loads / stores / waits-relevant ops, to check issue order.
usage: python tools/sass_timeline.py <object.o> <kernel-name-substring>"""
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
cur, keep = None, []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        keep.append(line)
ops = []
for l in keep:
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+)", l)
    if m:
        ops.append(m.group(3))
print("instructions:", len(ops))


def cls(o):
    if o.startswith("LDG"):
        if "CONSTANT" in o:
            return "Lc"
        if "STRONG.GPU" in o:
            return "Lg"
        if "EF" in o:
            return "Ls"
        return "L"
    if o.startswith("STG"):
        return "S"
    if o.startswith("POPC"):
        return "p"
    if o.startswith(("DADD", "DMUL", "DFMA", "DSETP")):
        return "d"
    if o.startswith("BRA"):
        return "|"
    if o.startswith("BAR"):
        return "B"
    return "."


s = "".join(cls(o) if len(cls(o)) == 1 else "<" + cls(o) + ">" for o in ops)
# collapse runs of '.' and 'd'
s = re.sub(r"\.+", lambda m: "." if len(m.group()) < 6 else f"[{len(m.group())}]", s)
s = re.sub(r"d{4,}", lambda m: f"(d{len(m.group())})", s)
print(s)
