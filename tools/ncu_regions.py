"""Per-routine share of the stall samples / executed instructions of one `ncu --set full --import-source on` capture of dm_step_kernel:
regions are cut at the routine definitions found in the captured source of dm_step.cu.  usage: python tools/ncu_regions.py <file.ncu-rep> [dm_step.cu of that build]"""
import csv, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = None; hdr = None; data = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 2 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r)); key = (cur, int(r[0]))
        if key in data: continue
        data[key] = (int(d['Instructions Executed']), int(d['# Samples']), r[1], int(d['stall_barrier']))
marks = []
srcfile = sys.argv[2] if len(sys.argv) > 2 else 'deepmimic_b200/csrc/kernels/dm_step.cu'   # the dm_step.cu the capture was built from (e.g. from `git show <commit>:...`)
for l, text in enumerate(open(srcfile).read().splitlines(), 1):
    for name in ('pgs_sweeps(', 'void solve_rows(', 'void kin_pass(', 'int collide(', 'float3 aba_solve(', 'float3 dv_pass(', 'void vel_pass(', 'dm_step_kernel(', 'V3 tile_com('):
        if name in text and ('__device__' in text or '__global__' in text or 'void pgs_sweeps' in text): marks.append((l, name.strip('( ').split()[-1]))
marks.sort()
tot_i = sum(v[0] for v in data.values()); tot_s = sum(v[1] for v in data.values()); tot_b = sum(v[3] for v in data.values())
agg = {}
for (f, l), (i, s, text, b) in data.items():
    if f != 'dm_step.cu': name = 'other files (dm_math.cuh, intrinsics: inlined helpers)'
    else:
        name = 'header'
        for ml, mn in marks:
            if l >= ml: name = mn
    a = agg.setdefault(name, [0, 0, 0]); a[0] += i; a[1] += s; a[2] += b
print('region                                             inst%   samples%  (of which barrier%)')
for name, (i, s, b) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print('%-50s %5.1f   %5.1f   %5.1f' % (name, 100.0 * i / tot_i, 100.0 * s / tot_s, 100.0 * b / tot_s))
