"""Aggregate warp-stall samples per CUDA source line from an .ncu-rep (needs -lineinfo + --import-source on).
    python tools/ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys, io
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = "?"
hdr = None
data = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[2] != "-": continue           # per-SASS rows carry an address; the per-line summary rows have "-"
    try:
        d = dict(zip(hdr[4:], r[4:]))
        data.append((int(d["# Samples"]), int(d["Instructions Executed"]), fname, r[0], r[1],
                     {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}))
    except Exception:
        pass
tot = sum(x[0] for x in data) or 1
print(f"total samples {tot}")
for s, ie, f, ln, src, st in sorted(data, key=lambda x: -x[0])[:top]:
    tops = ",".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{s:6d} {100*s/tot:5.1f}% inst={ie:7d} {f}:{ln}: {src.strip()[:90]}   [{tops}]")
