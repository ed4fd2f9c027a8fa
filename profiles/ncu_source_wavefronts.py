import csv,sys
def load(p):
    rows=list(csv.reader(open(p)))
    hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
    data=[r for r in rows[2:] if r and r[0].startswith('0x')]
    seen=set(); out=[]
    for r in data:
        if r[0] in seen: continue
        seen.add(r[0]); out.append(r)
    return ix,out
def I(r,ix,k):
    try: return int(r[ix[k]])
    except: return 0
ix,d=load(sys.argv[1])
nb=float(sys.argv[2]) if len(sys.argv)>2 else 1.0
tot=0
print('%-62s %9s %9s %9s %7s %8s'%('instr','exec/b','wf/b','ideal/b','smpl','long/short'))
for r in d:
    w=I(r,ix,'L1 Wavefronts Shared')
    tot+=w
    if w>0.2*nb or I(r,ix,'# Samples')>300:
        print('%-62s %9.2f %9.2f %9.2f %7d %5s/%5s'%(r[ix['Source']].strip()[:62], I(r,ix,'Instructions Executed')/nb, w/nb, I(r,ix,'L1 Wavefronts Shared Ideal')/nb, I(r,ix,'# Samples'), r[ix['stall_long_sb']], r[ix['stall_short_sb']]))
print('total smem wf/batch', tot/nb, 'total samples', sum(I(r,ix,'# Samples') for r in d), 'instr/batch', sum(I(r,ix,'Instructions Executed') for r in d)/nb)
