import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
N=int(sys.argv[2]) if len(sys.argv)>2 else 40
hdr=None; agg=collections.OrderedDict(); cur_file=None
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No' or r[0]=='Address': hdr=r; continue
    if hdr is None or hdr[0]!='Line No': continue
    d={}
    for h,v in zip(hdr,r):
        if h not in d: d[h]=v
    if not d.get('Line No','').isdigit(): continue
    try:
        ie=float(d['Instructions Executed'] or 0); te=float(d['Thread Instructions Executed'] or 0); s=float(d['# Samples'] or 0)
    except: continue
    if ie==0 and s==0: continue
    k=(cur_file,int(d['Line No']))
    o=agg.get(k,(0,0,0,''))
    agg[k]=(o[0]+ie,o[1]+te,o[2]+s,d['Source'].strip()[:95])
tot_inst=sum(v[0] for v in agg.values()); tot_s=sum(v[2] for v in agg.values())
print('total inst %.3g samples %d'%(tot_inst,tot_s))
key=(lambda kv:-kv[1][2]) if len(sys.argv)>3 and sys.argv[3]=='s' else (lambda kv:-kv[1][0])
for k,v in sorted(agg.items(), key=key)[:N]:
    print('%-14s %4d inst %5.1f%% thr %4.1f smp %5.1f%%  %s'%(k[0][:14],k[1],100*v[0]/tot_inst, v[1]/max(v[0],1), 100*v[2]/tot_s, v[3]))
