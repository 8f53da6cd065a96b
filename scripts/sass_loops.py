"""Instruction mix of the FFMA2-bearing loops of one correlate_kernel instantiation.
    python scripts/sass_loops.py 16 3 0"""
import collections, re, subprocess, sys
A, L, F = sys.argv[1:4]
tag = f"Li{A}ELi{L}ELb{F}"
sass = subprocess.run(["cuobjdump", "-sass", "gpuacceleratedtracking_b200/libgat.so"], capture_output=True, text=True).stdout
keep, on = [], False
for line in sass.split("\n"):
    if "Function :" in line:
        on = tag in line and "correlate_kernel" in line
    if on:
        keep.append(line)
ins = [l for l in keep if re.search(r"/\*[0-9a-f]{4}\*/\s", l)]
addr = lambda l: int(re.search(r"/\*([0-9a-f]{4})\*/", l).group(1), 16)
def op(l):
    m = re.search(r"\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    return m.group(2) if m else "?"
index = {addr(l): i for i, l in enumerate(ins)}
for i, l in enumerate(ins):
    m = re.search(r"BRA\S*\s+(?:U?P?\w+,\s*)?0x([0-9a-f]+)", l)
    if m and int(m.group(1), 16) < addr(l) and int(m.group(1), 16) in index:
        body = ins[index[int(m.group(1), 16)]:i + 1]
        c = collections.Counter(op(x).split(".")[0] for x in body)
        if c.get("FFMA2", 0) + c.get("FFMA", 0) > 4 and len(body) < 400:
            print(f"{tag} loop {addr(body[0]):#x}: {len(body)} instrs", dict(c.most_common(16)))
            if "--dump" in sys.argv:
                print("\n".join(x[:100] for x in body))
