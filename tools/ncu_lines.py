"""Join an ncu SASS source-page CSV with nvdisasm --print-line-info output to get per-source-line hot spots.
usage: ncu_lines.py <src.csv> <dis_line.txt> <mangled-kernel-substring> [topN]"""
import collections
import csv
import glob
import re
import sys

src_csv, dis, kern = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > 10]
num = lambda x: float(x) if x not in ('', 'N/A') else 0.0
# parse disassembly of the kernel: sequence of (file,line) per instruction
lines = open(dis).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.') and kern in l][0]
cur = ('?', 0)
seq = []
for l in lines[start + 1:]:
    if l.startswith('//---') or (l.startswith('.text.') and kern not in l):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        seq.append(cur)
print('instructions: ncu', len(data), 'nvdisasm', len(seq))
n = min(len(data), len(seq))
I, S = idx['Instructions Executed'], idx['# Samples']
stall_cols = [(h[6:], i) for h, i in idx.items() if h.startswith('stall_') and 'Not Issued' not in h]
tot_i = sum(num(r[I]) for r in data)
tot_s = sum(num(r[S]) for r in data)
agg = collections.defaultdict(lambda: [0.0, 0.0, 0, collections.Counter()])
tot_stall = collections.Counter()
for k in range(n):
    a = agg[seq[k]]
    a[0] += num(data[k][I])
    a[1] += num(data[k][S])
    a[2] += 1
    for nm, ci in stall_cols:
        v = num(data[k][ci])
        a[3][nm] += v
        tot_stall[nm] += v
srccache = {}


def text(f, ln):
    if f not in srccache:
        c = glob.glob('/root/repo/radex_emcee_b200/csrc/' + f)
        srccache[f] = open(c[0]).read().split('\n') if c else []
    t = srccache[f]
    return t[ln - 1].strip()[:80] if 0 < ln <= len(t) else ''


print('total samples %d, instructions executed %d' % (tot_s, tot_i))
print('stall totals:', ' '.join('%s=%.1f%%' % (k, 100 * v / max(1, tot_s)) for k, v in tot_stall.most_common(9)))
print('%6s %6s %5s  %s' % ('samp%', 'inst%', 'sass', 'location'))
for (f, ln), (i, s, c, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    top = ' '.join('%s=%.0f%%' % (k, 100 * v / max(1, s)) for k, v in st.most_common(3))
    print('%5.1f%% %5.1f%% %5d  %s:%d  %s   [%s]' % (100 * s / tot_s, 100 * i / tot_i, c, f, ln, text(f, ln), top))

# ---- coarse regions: functions of lvg_v2.cuh, and inside solve() the blocks introduced by "// ----" comments
src = srccache.get('lvg_v2.cuh') or (open(glob.glob('/root/repo/radex_emcee_b200/csrc/lvg_v2.cuh')[0]).read().split('\n'))
marks = []
for i, t in enumerate(src, 1):
    st = t.strip()
    if st.startswith('__device__') and '~' not in st and not st.startswith('__device__ unsigned long long g_t'):
        m = re.search(r'(\w+)\s*\(', st.split('__forceinline__')[-1].split('__noinline__')[-1])
        marks.append((i, m.group(1) if m else st[:30]))
    elif st.startswith('// ----') and marks and marks[-1][1].startswith('solve'):
        marks.append((i, 'solve: ' + st[7:60].strip(' -')))
reg = collections.defaultdict(lambda: [0.0, 0.0])
for (f, ln), (i, s_, c, st) in agg.items():
    name = f
    if f == 'lvg_v2.cuh':
        name = 'lvg_v2.cuh:?'
        for a, nm in marks:
            if a <= ln:
                name = nm
    reg[name][0] += i
    reg[name][1] += s_
print('\nregions (inst%, samp%):')
for nm, (i, s_) in sorted(reg.items(), key=lambda kv: -kv[1][1]):
    if s_ / tot_s > 0.002:
        print('%5.1f%% %5.1f%%  %s' % (100 * i / tot_i, 100 * s_ / tot_s, nm))
