import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; data=rows[hi+1:]
ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value')
agg=collections.defaultdict(lambda: collections.defaultdict(list))
for r in data:
    if len(r)<=iv: continue
    k=r[ik].split('(')[0].replace('void ','')
    try: agg[k][r[im]].append(float(r[iv].replace(',','')))
    except: pass
tot=0
for k in sorted(agg):
    d=agg[k]
    t=d['gpu__time_duration.sum']; tot+=sum(t)
    f=lambda x: sum(d[x])/max(len(d[x]),1)
    print('%-42s n=%3d %7.1f us/launch sum %6.2f ms inst %.2e warps_act %4.1f%% issue %4.1f%% | no_inst %.2f long_sb %.2f wait %.2f short_sb %.2f branch %.2f thr/inst %.1f'%(k,len(t),sum(t)/len(t)/1e3,sum(t)/1e6,f('smsp__inst_executed.sum'),f('sm__warps_active.avg.pct_of_peak_sustained_active'),f('smsp__issue_active.avg.pct_of_peak_sustained_active'),f('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio'),f('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'),f('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio'),f('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio'),f('smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio'),f('smsp__thread_inst_executed_per_inst_executed.ratio')))
print('total ms', tot/1e6)
