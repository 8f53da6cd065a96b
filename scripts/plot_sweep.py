"""JSONL of `bench.py --sweep` -> the reference paper's figure: processing time of one 1 ms correlate call against the
sampling frequency, log-log, one panel per (GNSS, antennas, correlators), with the real-time line at 1 ms.

    python bench.py --sweep > sweep.jsonl          (on a B200 box)
    python scripts/plot_sweep.py sweep.jsonl out.svg [out.md]

Follows /root/reference/scripts/plot_benchmarks.jl (panels GPS L1 1/3, 4/3, 4/7 antennas/correlators + GPS L5; x = sampling
frequency, y = processing time, log10 axes) and src/plots.jl.  The reference plots one curve per algorithm id; here the curves
are the CPU port (1 thread), the synchronous GPU call and the back-to-back GPU device time.  No plotting package is
installed in this image (matplotlib, Makie ...), so the SVG is written by hand; the optional third argument gets the same
numbers as a Markdown table."""
import json
import math
import sys

CURVES = [("CPU_ns", "CPU port, 1 thread", "#d62728"), ("GPU_sync_ns", "B200 call + sync", "#1f77b4"),
          ("GPU_resident_ns", "B200 resident session", "#9467bd"), ("GPU_device_ns", "B200 back-to-back", "#2ca02c")]
W, H, ML, MB, MT, MR = 330, 250, 52, 40, 26, 10


def log_ticks(lo, hi):
    return [10.0 ** e for e in range(math.floor(math.log10(lo)), math.ceil(math.log10(hi)) + 1)]


def panel(rows, title, x0, y0):
    xs = [r["sampling_frequency_hz"] for r in rows]
    ys = [r[k] * 1e-9 for r in rows for k, _, _ in CURVES if r.get(k, 0) > 0]
    xlo, xhi = min(xs) * 0.8, max(xs) * 1.25
    ylo, yhi = min(min(ys) * 0.7, 1e-6), max(max(ys) * 1.4, 2e-3)
    px = lambda x: x0 + ML + (math.log10(x) - math.log10(xlo)) / (math.log10(xhi) - math.log10(xlo)) * (W - ML - MR)
    py = lambda y: y0 + H - MB - (math.log10(y) - math.log10(ylo)) / (math.log10(yhi) - math.log10(ylo)) * (H - MB - MT)
    out = [f'<rect x="{x0 + ML}" y="{y0 + MT}" width="{W - ML - MR}" height="{H - MB - MT}" fill="none" stroke="#444"/>',
           f'<text x="{x0 + W / 2}" y="{y0 + 16}" text-anchor="middle" font-size="12">{title}</text>']
    for t in log_ticks(xlo, xhi):
        if xlo <= t <= xhi:
            out.append(f'<line x1="{px(t):.1f}" y1="{y0 + MT}" x2="{px(t):.1f}" y2="{y0 + H - MB}" stroke="#ddd"/>')
            out.append(f'<text x="{px(t):.1f}" y="{y0 + H - MB + 14}" text-anchor="middle" font-size="9">1e{int(round(math.log10(t)))}</text>')
    for t in log_ticks(ylo, yhi):
        if ylo <= t <= yhi:
            out.append(f'<line x1="{x0 + ML}" y1="{py(t):.1f}" x2="{x0 + W - MR}" y2="{py(t):.1f}" stroke="#ddd"/>')
            out.append(f'<text x="{x0 + ML - 4}" y="{py(t) + 3:.1f}" text-anchor="end" font-size="9">1e{int(round(math.log10(t)))}</text>')
    out.append(f'<line x1="{x0 + ML}" y1="{py(1e-3):.1f}" x2="{x0 + W - MR}" y2="{py(1e-3):.1f}" stroke="#000" stroke-dasharray="5,3"/>')
    out.append(f'<text x="{x0 + W - MR - 3}" y="{py(1e-3) - 3:.1f}" text-anchor="end" font-size="9">real time (1 ms)</text>')
    for key, _, colour in CURVES:
        pts = [(px(r["sampling_frequency_hz"]), py(r[key] * 1e-9)) for r in rows if r.get(key, 0) > 0]   # (0 = not measured)
        if not pts:
            continue
        out.append('<polyline fill="none" stroke="%s" stroke-width="1.6" points="%s"/>' % (colour, " ".join(f"{a:.1f},{b:.1f}" for a, b in pts)))
        out += [f'<circle cx="{a:.1f}" cy="{b:.1f}" r="2.2" fill="{colour}"/>' for a, b in pts]
    out.append(f'<text x="{x0 + W / 2}" y="{y0 + H - 6}" text-anchor="middle" font-size="10">sampling frequency [Hz]</text>')
    out.append(f'<text transform="translate({x0 + 11},{y0 + H / 2}) rotate(-90)" text-anchor="middle" font-size="10">processing time [s]</text>')
    return out


def main():
    src, dst = sys.argv[1], sys.argv[2]
    meta, rows = {}, []
    for line in open(src):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        if "metadata" in d:
            meta = d["metadata"]
        elif "num_samples" in d:
            rows.append(d)
    groups = {}
    for r in rows:
        groups.setdefault((r["system"], r["num_ants"], r["num_correlators"]), []).append(r)
    keys = sorted(groups, key=lambda k: (k[0], k[1], k[2]))
    cols = 3
    n_rows = (len(keys) + cols - 1) // cols
    svg = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{cols * W}" height="{n_rows * H + 40}" font-family="sans-serif">',
           f'<rect width="100%" height="100%" fill="white"/>']
    for i, k in enumerate(keys):
        g = sorted(groups[k], key=lambda r: r["num_samples"])
        name = {"GPSL1": "GPS L1 C/A", "GPSL5": "GPS L5"}.get(k[0], k[0])
        svg += panel(g, f"{name}, {k[1]} antenna(s), {k[2]} correlators", (i % cols) * W, (i // cols) * H)
    lx = 10
    for _, label, colour in CURVES:
        svg.append(f'<line x1="{lx}" y1="{n_rows * H + 18}" x2="{lx + 22}" y2="{n_rows * H + 18}" stroke="{colour}" stroke-width="2"/>')
        svg.append(f'<text x="{lx + 26}" y="{n_rows * H + 22}" font-size="11">{label}</text>')
        lx += 175
    svg.append(f'<text x="{lx}" y="{n_rows * H + 22}" font-size="10" fill="#555">{meta.get("GPU_model", "")} / {meta.get("CPU_model", "")}; estimator: minimum</text>')
    svg.append("</svg>")
    open(dst, "w").write("\n".join(svg) + "\n")
    if len(sys.argv) > 3:
        with open(sys.argv[3], "w") as f:
            f.write("| system | antennas | correlators | samples | CPU 1 thread [us] | GPU call + sync [us] | GPU resident session [us] | GPU back-to-back [us] | real time |\n|---|---|---|---|---|---|---|---|---|\n")
            for k in keys:
                for r in sorted(groups[k], key=lambda r: r["num_samples"]):
                    res = f"{r['GPU_resident_ns'] / 1e3:.1f}" if r.get("GPU_resident_ns", 0) > 0 else "-"
                    f.write(f"| {k[0]} | {k[1]} | {k[2]} | {r['num_samples']} | {r['CPU_ns'] / 1e3:.1f} | {r['GPU_sync_ns'] / 1e3:.1f} | {res} | "
                            f"{r['GPU_device_ns'] / 1e3:.1f} | {'yes' if r['realtime'] else 'no'} |\n")
    print(f"{len(rows)} points in {len(keys)} panels -> {dst}")


if __name__ == "__main__":
    main()
