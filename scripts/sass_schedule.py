"""Static issue schedule of the FFMA2 loop of one correlate_kernel instantiation: decodes the per-instruction
control word (stall count, yield, scoreboard waits) from the SASS encodings.
    python scripts/sass_schedule.py 16 3 0 [sc16]"""
import os, re, subprocess, sys
pos = [a for a in sys.argv[1:] if not a.startswith("--")]
A, L, F = pos[:3]
SC = pos[3] if len(pos) > 3 else "0"
tag = f"Li{A}ELi{L}ELb{F}ELb{SC}E"
sass = subprocess.run(["cuobjdump", "-sass", os.environ.get("GAT_LIB_PATH", "gpuacceleratedtracking_b200/libgat.so")], capture_output=True, text=True).stdout
lines, on = [], False
for line in sass.split("\n"):
    if "Function :" in line:
        on = tag in line and "correlate_kernel" in line
    if on:
        lines.append(line)
# instruction line: /*addr*/ text ; /* 0xLOW */   followed by a line with /* 0xHIGH */
ins = []
for i, l in enumerate(lines):
    m = re.search(r"/\*([0-9a-f]{4})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", l)
    if m and i + 1 < len(lines):
        h = re.search(r"/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
        if h:
            hi = int(h.group(1), 16)
            ctrl = hi >> 41            # 23 control bits: [3:0] stall, [4] yield, [7:5] wr bar, [10:8] rd bar, [16:11] wait mask, [20:17] reuse
            ins.append((int(m.group(1), 16), m.group(2).strip(), ctrl & 0xf, (ctrl >> 4) & 1, (ctrl >> 5) & 7, (ctrl >> 8) & 7, (ctrl >> 11) & 0x3f))
index = {a: i for i, (a, *_r) in enumerate(ins)}
for i, (a, text, *_r) in enumerate(ins):
    m = re.search(r"BRA\S*\s+(?:U?P?\w+,\s*)?0x([0-9a-f]+)", text)
    if m and int(m.group(1), 16) < a and int(m.group(1), 16) in index:
        body = ins[index[int(m.group(1), 16)]:i + 1]
        nf = sum(1 for b in body if b[1].split()[0].startswith(("FFMA2", "FMUL2", "FADD2")) or (b[1].startswith("@") and "FFMA2" in b[1]))
        if nf > 4 and len(body) < 400:
            stalls = sum(b[2] for b in body)
            print(f"{tag} loop {body[0][0]:#x}: {len(body)} instrs, {nf} packed FP32, sum of stall counts {stalls} cycles "
                  f"(FMA-pipe floor {2 * nf})")
            if "--dump" in sys.argv:
                for a_, t_, st, y, wb, rb, wm in body:
                    print(f"  {a_:04x} st={st:2d} y={y} wr={wb if wb != 7 else '-'} rd={rb if rb != 7 else '-'} wait={wm:06b}  {t_[:90]}")
