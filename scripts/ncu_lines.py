"""Summarise an ncu `--page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[2]
ci = {h: i for i, h in enumerate(hdr)}
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
iI, iS, iT = ci["Instructions Executed"], ci["# Samples"], ci["Thread Instructions Executed"]
out = []
for r in rows[3:]:
    if len(r) < len(hdr) or not r[0]:
        continue
    try:
        out.append((int(r[0]), r[1], float(r[iI]), float(r[iS]), float(r[iT])))
    except ValueError:
        pass
totI = sum(o[2] for o in out)
totS = sum(o[3] for o in out)
print(f"total inst {totI:.0f} ({totI/steps:.1f}/step)  samples {totS:.0f}")
out.sort(key=lambda o: -o[3])
print("--- top lines by stall samples")
for ln, src, ins, smp, thr in out[:45]:
    print(f"{ln:4d} inst/step={ins/steps:7.1f} samp%={100*smp/totS:5.1f} thr/inst={thr/max(ins,1):5.1f} | {src[:110]}")
print("--- top lines by instructions")
out.sort(key=lambda o: -o[2])
for ln, src, ins, smp, thr in out[:40]:
    print(f"{ln:4d} inst/step={ins/steps:7.1f} samp%={100*smp/totS:5.1f} thr/inst={thr/max(ins,1):5.1f} | {src[:110]}")
