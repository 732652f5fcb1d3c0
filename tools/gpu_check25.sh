mkdir -p gpurun_out
LVD_GN_NO_FUSED=1 PERF_NORM_EAGER=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"gn_|layernorm" -c 400 --csv --log-file gpurun_out/ncu_norm25.csv python tools/perf_norm.py > gpurun_out/ncu25.log 2>&1; tail -3 gpurun_out/ncu25.log
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/ncu_norm25.csv')) if len(r)>10]
hdr=rows[0]; idx={h:i for i,h in enumerate(hdr)}
agg=collections.OrderedDict()
for r in rows[1:]:
    k=(r[idx['Kernel Name']][:40], r[idx['Grid Size']], r[idx['Block Size']])
    agg.setdefault(k,{}).setdefault(r[idx['Metric Name']],[]).append(float(r[idx['Metric Value']].replace(',','')))
for k,v in agg.items():
    t=sorted(v['gpu__time_duration.sum'])[len(v['gpu__time_duration.sum'])//2]
    rd=sum(v['dram__bytes_read.sum'])/len(v['dram__bytes_read.sum']); wr=sum(v['dram__bytes_write.sum'])/len(v['dram__bytes_write.sum'])
    print(k, 'n=%d'%len(v['gpu__time_duration.sum']), 't=%.1f'%t, 'rd=%.1f wr=%.1f'%(rd,wr), 'warps=%.0f%%'%(sum(v['sm__warps_active.avg.pct_of_peak_sustained_active'])/len(v['sm__warps_active.avg.pct_of_peak_sustained_active'])))
PY
