"""Opcode mix (dynamic, warp-level) and static SASS size per source line from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys, collections, re
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ops = collections.Counter(); stat_line = collections.Counter(); dyn_line = collections.Counter(); stall_op = collections.Counter()
hdr = None; cur = None; cur_file = ""
def _i(x):
    try: return int(x)
    except (ValueError, TypeError): return 0
for r in csv.reader(open(path, newline="")):
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[0] != "":
        cur = (cur_file, _i(r[0])); continue
    sass = r[3]; ex = _i(r[7]); smp = _i(r[6])
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", sass.strip())
    op = m.group(2) if m else "?"
    ops[op] += ex; stall_op[op] += smp
    stat_line[cur] += 1; dyn_line[cur] += ex
tot = sum(ops.values()); ts = sum(stall_op.values())
print("total warp instructions", tot, "static SASS", sum(stat_line.values()))
for op, c in ops.most_common(top): print(f"{op:12s} {c:12d} {100*c/tot:5.1f}%   samples {100*stall_op[op]/max(ts,1):5.1f}%")
print("---- static SASS instructions per source line")
for k, c in stat_line.most_common(top): print(f"{c:6d} static  {dyn_line[k]:12d} dyn  {k[0]}:{k[1]}")
