"""Summarise an .ncu-rep (ncu --set full --import-source on) as text: duration, instruction / pipe / shared-memory counters,
stall breakdown and the executed-instruction mix by opcode.  Usage: python tools/ncu_summary.py report.ncu-rep
(also writes the SASS source page to /tmp/sass/last_src.csv for tools/ncu_segments.py)."""
import os
os.makedirs("/tmp/sass", exist_ok=True)
import csv, collections, subprocess, sys
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; d=dict(zip(hdr,rows[2]))
def g(k): return d.get(k,'?')
for k in ['gpu__time_duration.sum','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_active','launch__registers_per_thread','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active']:
    print(k, g(k))
st={k:float(v.replace(',','')) for k,v in d.items() if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k and v not in ('','n/a')}
tot=sum(st.values())
print('stalls:', ', '.join(f'{k.split("stalled_")[1]} {100*v/tot:.1f}%' for k,v in sorted(st.items(), key=lambda x:-x[1])[:9]))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; data=rows[2:]
iA=hdr.index('Source'); iE=hdr.index('Instructions Executed'); iS=hdr.index('# Samples')
byop=collections.Counter(); sm=collections.Counter(); tot=0
for r in data:
    try: e=int(r[iE]); s=int(r[iS])
    except: continue
    t=r[iA].split(); op=t[1] if t[0].startswith('@') else t[0]; op=op.split('.')[0]
    byop[op]+=e; sm[op]+=s; tot+=e
ts=sum(sm.values())
print('total',tot)
for op,c in byop.most_common(16): print(f'{op:10s} {c/1e6:8.2f}M {100*c/tot:5.1f}%  samples {100*sm[op]/ts:5.1f}%')
open('/tmp/sass/last_src.csv','w').write(src)
