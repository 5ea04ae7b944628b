"""Attribute ncu source-page samples / instructions of a flux kernel to its stages.
usage: python dev/stage_attrib.py <rep> <kernel-substring> <csv-cache>"""
import csv, subprocess, sys, os
rep, kname = sys.argv[1], sys.argv[2]
cache = sys.argv[3] if len(sys.argv) > 3 else '/tmp/src.csv'
if not os.path.isfile(cache):
    open(cache, 'w').write(subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'sass,cuda', '--csv'],
                                          capture_output=True, text=True).stdout)
rows = list(csv.reader(open(cache)))
cur = fn = None; hdr = None; data = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name': fn = r[1]; continue
    if len(r) > 5 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) > 7 and r[0].isdigit() and kname in fn:
        s = int(r[4]) if r[4].isdigit() else 0; i = int(r[7]) if r[7].isdigit() else 0
        k = (cur, int(r[0]))
        o = data.get(k, (0, 0, r[1].strip()[:100]))
        data[k] = (o[0] + s, o[1] + i, o[2])
tot_s = sum(v[0] for v in data.values()); tot_i = sum(v[1] for v in data.values())
print('kernel', kname, 'samples', tot_s, 'warp instructions', tot_i)
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'xpsi_b200/csrc/integrate_azinv.cu')).read().split('\n')
def find(pat, start=0):
    for n, l in enumerate(src[start:], start + 1):
        if pat in l: return n
    return None
F = 'integrate_azinv.cu'
marks = [(l, n) for n, l in enumerate(src, 1)]
def agg(lo, hi, name):
    s = sum(v[0] for k, v in data.items() if k[0] == F and lo <= k[1] <= hi)
    i = sum(v[1] for k, v in data.items() if k[0] == F and lo <= k[1] <= hi)
    print('%-44s lines %4d-%4d samples %5.1f%%  inst %5.1f%%' % (name, lo, hi, 100 * s / tot_s, 100 * i / tot_i))
if 'mma' in kname:
    k0 = find('k_azinv_flux_mma(AzinvArgs a')
    pts = [('setup + image head', k0), ('stage 1 leaf profile', find('// ---- (1) leaf profile: thread = leaf', k0)),
           ('stage 2 coefficients', find('// ---- (2) cubic pieces + positivity flags: thread', k0)),
           ('stage 3 DMMA sweep', find('// ---- (3) tiles x cubic pieces on the tensor cores', k0)),
           ('stage 3 slow pass', find('if (anyf != 0ull) {', k0)),
           ('final store', find('// ---- ring/chunk contribution -> flux[q, e, k] (RED) or its slot', k0)),
           ('end', find('// deterministic mode: flux[q, e, k] = sum over the member', k0))]
else:
    k0 = find('k_azinv_flux(AzinvArgs a')
    pts = [('setup + image head', k0), ('stage 1 leaf profile', find('// ---- (1) leaf profile (pyx:445-478)', k0)),
           ('stage 2 coefficients', find('// ---- (2) phase-spline coefficients + positivity flags', k0)),
           ('stage 3 accumulation', find('// ---- (3) interval moments x spline coefficients', k0)),
           ('final', find('// ---- ring/chunk contribution -> flux[q, e, k] ---', k0)),
           ('end', find('// ----------------------------------------------------------------------------------------------', k0))]
for (n, lo), (_, hi) in zip(pts[:-1], pts[1:]):
    agg(lo, hi - 1, n)
lo, hi = pts[0][1], pts[-1][1]
oth = sorted([(k, v) for k, v in data.items() if not (k[0] == F and lo <= k[1] < hi)], key=lambda kv: -kv[1][0])
so = sum(v[0] for k, v in oth); io = sum(v[1] for k, v in oth)
print('inlined helpers outside the kernel body: samples %.1f%% inst %.1f%%' % (100 * so / tot_s, 100 * io / tot_i))
for k, v in oth[:14]:
    print('  %5.1f%% inst %5.1f%% %s:%d %s' % (100 * v[0] / tot_s, 100 * v[1] / tot_i, k[0], k[1], v[2]))
print('top kernel-body lines')
ins = sorted([(k, v) for k, v in data.items() if k[0] == F and lo <= k[1] < hi], key=lambda kv: -kv[1][0])[:22]
for k, v in ins:
    print('  %5.1f%% inst %5.1f%% :%d %s' % (100 * v[0] / tot_s, 100 * v[1] / tot_i, k[1], v[2]))
