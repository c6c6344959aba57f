import sys, time, numpy as np
sys.path.insert(0, '.')
import sibelia_b200 as sb
def read_fasta(p):
    seqs=[]; cur=[]
    for line in open(p,'rb'):
        if line.startswith(b'>'):
            if cur: seqs.append(b''.join(cur)); cur=[]
        else: cur.append(line.strip().upper())
    if cur: seqs.append(b''.join(cur))
    return seqs
seqs = read_fasta('oracle/_ref/data/Staphylococcus.fasta')
rng = np.random.default_rng(0)
chrs=[]
for s in seqs:
    a = np.frombuffer(s, dtype=np.uint8).copy()
    bad = ~np.isin(a, np.frombuffer(b'ACGT', dtype=np.uint8))
    a[bad] = np.frombuffer(b'ACGT', dtype=np.uint8)[rng.integers(0,4,int(bad.sum()))]
    chrs.append(a)
print([len(c) for c in chrs])
ctx = sb.Context(0)
ctx.upload(chrs)
ctx.set_profiling(True)
for k in (30, 100, 100, 1000, 5000):
    t=time.perf_counter(); c, n = ctx.enumerate_resident(k); dt=(time.perf_counter()-t)*1e3
    st = sorted(ctx.kernel_stats(), key=lambda x:-x['ms'])[:6]
    print("k=%d V=%d I=%d wall %.2f ms device %.2f ms" % (k, c, n, dt, ctx.last_device_ms()), [(s['name'], s['launches'], round(s['ms'],3)) for s in st])
