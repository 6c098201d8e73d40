"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by source line: samples and top stall reasons."""
import collections
import csv
import sys


def num(x):
    try:
        return int(float(x))
    except Exception:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 45
cur_file = None
idx = {}
agg = collections.Counter()
inst = collections.Counter()
stall = collections.defaultdict(collections.Counter)
src = {}
key = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        idx = {n: i for i, n in enumerate(r)}
        continue
    if r[0] != '':
        try:
            key = (cur_file, int(r[0]))
        except Exception:
            continue
        src[key] = r[1]
        continue
    if key is None:
        continue
    agg[key] += num(r[idx['# Samples']])
    for n in ('Instructions Executed', 'inst_executed'):
        if n in idx:
            inst[key] += num(r[idx[n]])
            break
    for n in ('stall_barrier', 'stall_long_sb', 'stall_no_inst', 'stall_wait', 'stall_short_sb', 'stall_math', 'stall_lg', 'stall_branch_resolving', 'stall_mio', 'stall_membar'):
        if n in idx:
            stall[key][n] += num(r[idx[n]])
tot = sum(agg.values())
print('columns:', sorted(idx))
print('total samples', tot, 'total warp instructions executed', sum(inst.values()))
if inst:
    ti = sum(inst.values())
    print('--- by executed warp instructions')
    for key, s in inst.most_common(top_n):
        print(f"{key[0]}:{key[1]:5d} {s:12d} {100 * s / ti:5.1f}%  {src[key][:110]}")
    byfile = collections.Counter()
    for key, s in inst.items():
        byfile[key[0]] += s
    print('--- by file', {k: f"{100 * v / ti:.1f}%" for k, v in byfile.most_common()})
    print('--- by stall samples')
for key, s in agg.most_common(top_n):
    st = stall[key]
    top = ', '.join(f"{k[6:]}:{v}" for k, v in st.most_common(3) if v)
    print(f"{key[0]}:{key[1]:5d} {s:6d} {100 * s / tot:5.1f}%  [{top}]  {src[key][:100]}")
