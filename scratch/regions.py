import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ntiles = float(sys.argv[2]) if len(sys.argv)>2 else 6016
hdr = rows[1]; ix = {h:i for i,h in enumerate(hdr)}; data = rows[2:]
def f(r,k):
    try: return float(r[ix[k]])
    except: return 0.0
regions=[]; cur={'start':data[0][ix['Address']][-5:], 'exec':0,'samp':0,'n':0,'ops':{}}
for r in data:
    src=r[ix['Source']].strip()
    cur['exec']+=f(r,'Instructions Executed'); cur['samp']+=f(r,'# Samples'); cur['n']+=1
    op=src.split()[1] if src.startswith('@') else src.split()[0]
    op=op.split('.')[0]
    cur['ops'][op]=cur['ops'].get(op,0)+f(r,'Instructions Executed')
    if src.startswith('BAR.SYNC') or 'RET' in src or src.startswith('EXIT'):
        regions.append(cur); cur={'start':r[ix['Address']][-5:], 'exec':0,'samp':0,'n':0,'ops':{}}
regions.append(cur)
tot=sum(x['exec'] for x in regions); ts=sum(x['samp'] for x in regions)
for x in regions:
    if x['exec']<1e5: continue
    top=sorted(x['ops'].items(), key=lambda kv:-kv[1])[:7]
    print(f"{x['start']} n={x['n']:4d} exec={x['exec']/ntiles:8.0f}/tile ({100*x['exec']/tot:4.1f}%) samp={100*x['samp']/ts:5.1f}%  " + ' '.join(f"{k}:{v/ntiles:.0f}" for k,v in top))
print('total per tile', tot/ntiles)
