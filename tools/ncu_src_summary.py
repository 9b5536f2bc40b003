"""Summarise `ncu --page source --csv --print-source cuda,sass` output: stall samples per CUDA source line, per kernel."""
import csv, sys, collections, re
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kern = None; hdr = None; mode = None
agg = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0, ""]))   # kernel -> line -> [samples, insts, src]
for r in csv.reader(open(path, newline="")):
    if not r: continue
    if r[0] == "Function Name": kern = re.sub(r"\(.*", "", r[1]); hdr = None; continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Line No" or r[0] == "Address": hdr = r; continue
    if hdr is None or kern is None: continue
    d = dict(zip(hdr, r))
    if "Line No" in d and hdr[0] == "Line No":
        try: ln = int(d["Line No"])
        except ValueError: continue
        def _i(x):
            try: return int(x)
            except (ValueError, TypeError): return 0
        s = _i(d.get("# Samples", "0")); ins = _i(d.get("Instructions Executed", "0"))
        a = agg[kern][(cur_file.split("/")[-1], ln)]
        a[0] += s; a[1] += ins; a[2] = r[1][:110]
for k, lines in agg.items():
    tot = sum(v[0] for v in lines.values())
    if tot == 0: continue
    print("=====", k, "total samples", tot)
    for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{v[0]:6d} {100*v[0]/tot:5.1f}%  inst={v[1]:8d}  {f}:{ln}  {v[2].strip()}")
