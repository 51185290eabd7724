import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals))
def g(k): 
    return d.get(k,'?')
print('duration us', g('gpu__time_duration.sum'), ' regs', g('launch__registers_per_thread'), ' inst', g('smsp__inst_executed.sum'))
print('ipc', g('sm__inst_executed.avg.per_cycle_active'), ' warps_active/sched', g('smsp__warps_active.avg.per_cycle_active'), ' eligible', g('smsp__warps_eligible.avg.per_cycle_active'))
print('dram rd MB', g('dram__bytes_read.sum'), ' wr', g('dram__bytes_write.sum'))
ks = [k for k in hdr if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k]
items=[]
for k in ks:
    try: items.append((float(d[k]),k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
    except: pass
print('stalls/issue:', ', '.join(f'{k}={v:.2f}' for v,k in sorted(items, reverse=True)[:9]))
print('shared bank conflicts ld/st', g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum'), g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum'))
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; ix = {h:i for i,h in enumerate(hdr)}; data = rows[2:]
def f(r,k):
    try: return float(r[ix[k]])
    except: return 0.0
tot = sum(f(r,'# Samples') for r in data)
print('n instr', len(data), 'samples', tot)
n = int(sys.argv[2]) if len(sys.argv)>2 else 14
for r in sorted(data, key=lambda r:-f(r,'# Samples'))[:n]:
    st = {k: f(r,k) for k in hdr if k.startswith('stall_') and 'Not Issued' not in k}
    big = sorted(st.items(), key=lambda kv:-kv[1])[:2]
    print(f"{r[ix['Address']][-5:]} {100*f(r,'# Samples')/tot:5.1f}%  {r[ix['Source']][:56]:56s} {[(k[6:],int(v)) for k,v in big]}")
agg = {}
for r in data:
    for k in hdr:
        if k.startswith('stall_') and 'Not Issued' not in k: agg[k] = agg.get(k,0)+f(r,k)
print('stall samples:', ', '.join(f'{k[6:]}={int(v)}' for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:9]))
