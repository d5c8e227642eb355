mkdir -p gpurun_out
export RP_XCHG_DEBUG=1 RP_XCHG_PROBES=100000000
python tools/xchg_local_bench.py --k 13 --world 2 --reads 100000 --whole > gpurun_out/xl_amb.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/xl_amb_launches.csv python tools/xchg_local_bench.py --k 13 --world 2 --reads 100000 --whole > gpurun_out/xl_amb_ncu.log 2>&1
tail -12 gpurun_out/xl_amb.log
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/xl_amb_launches.csv', errors='replace')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
for r in rows[hdr+1:]:
    if len(r)>vi:
        name=r[ki][:90]; v=float(r[vi].replace(',','')); u=r[ui]
        if u=='us': v/=1e3
        elif u=='ns': v/=1e6
        elif u=='s': v*=1e3
        if v>0.3: print('%9.2f ms  %s'%(v,name))
PY
