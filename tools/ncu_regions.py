"""Per-region (source line ranges of plba_warp.h etc.) instruction counts and stall-reason samples from the ncu source page CSV."""
import csv, collections, sys
path = sys.argv[1]
hdr = None; cur = None; cur_file = ''
reg = collections.defaultdict(collections.Counter)
def _i(x):
    try: return int(x)
    except Exception: return 0
import re
# regions are found by function markers in plba_warp.h: map line -> enclosing function by scanning the source
src = open('pl_slam_plucker_b200/csrc/plba_warp.h').read().split('\n')
marks = []
for i, l in enumerate(src, 1):
    m = re.search(r'(obs_lin_w|wtask_decode|wtask_accumulate|wtask_flush|assemble_items_w|update_items_w|k_assemble_w|k_update_w|struct LmLoad|orth_to_plk_sc)\b.*[{(]', l)
    if m and (l.startswith('PLBA_') or l.startswith('template') or l.startswith('struct') or 'PLBA_D void' in l or 'PLBA_HD' in l): marks.append((i, m.group(1)))
def region(f, ln):
    if f == 'plba_warp.h':
        name = 'warp.h head'
        for i, n in marks:
            if i <= ln: name = n
        if name in ('assemble_items_w', 'update_items_w'):
            txt = '\n'.join(src[max(0, ln - 40):ln])
            ph = [m.start() for m in re.finditer(r'// ---- ', txt)]
            if ph: name += ':' + txt[ph[-1] + 8:ph[-1] + 30].split('(')[0].strip()
        return name
    if f == 'plba_math.h': return 'math:inverse' if ln >= 330 else ('math:lines' if 202 <= ln <= 245 or 125 <= ln <= 137 else 'math:points/other')
    return f
for r in csv.reader(open(path, newline="")):
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[0] != "": cur = (cur_file, _i(r[0])); continue
    d = dict(zip(hdr, r)); g = region(*cur)
    c = reg[g]
    c['inst'] += _i(d['Instructions Executed']); c['samples'] += _i(d['# Samples']); c['static'] += 1
    for k in ('stall_wait', 'stall_no_inst', 'stall_long_sb', 'stall_short_sb', 'stall_math', 'stall_branch_resolving', 'stall_selected', 'stall_not_selected'): c[k] += _i(d.get(k, 0))
tot = sum(v['samples'] for v in reg.values()); ti = sum(v['inst'] for v in reg.values())
print(f"{'region':42s} {'Minst':>8s} {'inst%':>6s} {'smp%':>6s} {'static':>6s} | wait noinst longsb shortsb math branch sel notsel (% of region samples)")
for g, v in sorted(reg.items(), key=lambda kv: -kv[1]['samples']):
    s = max(v['samples'], 1)
    print(f"{g:42s} {v['inst']/1e6:8.1f} {100*v['inst']/ti:6.1f} {100*v['samples']/tot:6.1f} {v['static']:6d} | " + ' '.join(f"{100*v[k]/s:5.0f}" for k in ('stall_wait', 'stall_no_inst', 'stall_long_sb', 'stall_short_sb', 'stall_math', 'stall_branch_resolving', 'stall_selected', 'stall_not_selected')))
