"""Per source line / per kernel phase shares of executed instructions and stall samples from an ncu capture.

    ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv   (kernels built with -lineinfo)
    python profiles/source_hotspots.py src.csv [N]

Prints the totals, the share of every `// ---- phase X` section of lsq_filter_kernel and of every other source file,
then the N (default 30) source lines with the most stall samples.
"""
import csv,collections,sys,bisect,re
def load(path):
    rows=list(csv.reader(open(path)))
    cur=None; hdr=None
    agg=collections.defaultdict(lambda:[0,0,0]); src={}
    for r in rows:
        if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
        if len(r)>3 and r[0]=='Line No': hdr=r; continue
        if hdr and len(r)==len(hdr) and r[0].isdigit():
            i_inst=hdr.index('Instructions Executed'); i_s=hdr.index('# Samples'); i_t=hdr.index('Thread Instructions Executed')
            try:
                k=(cur,int(r[0])); agg[k][0]+=int(r[i_inst] or 0); agg[k][1]+=int(r[i_s] or 0); agg[k][2]+=int(r[i_t] or 0); src[k]=r[1]
            except ValueError: pass
    return agg,src
agg,src=load(sys.argv[1])
tot=sum(v[0] for v in agg.values()); ts=sum(v[1] for v in agg.values())
print('total inst %.3fG samples %d'%(tot/1e9,ts))
# phases by marker lines found in the embedded source
import os
ksrc=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),'lesseq_b200','csrc','lsq_kernels.cuh')
klines_txt=open(ksrc).read().split('\n')  # comment lines carry no SASS, so the markers come from the file on disk
marks=[(i+1,l.strip()) for i,l in enumerate(klines_txt) if '// ---- phase' in l]
ph=collections.defaultdict(lambda:[0,0])
fn=collections.defaultdict(lambda:[0,0])
klines=sorted(k[1] for k in src if k[0]=='lsq_kernels.cuh')
# find emit lambda range
emit0=min((i+1 for i,l in enumerate(klines_txt) if 'auto emit' in l),default=0)
emit1=min((i+1 for i,l in enumerate(klines_txt) if i+1>emit0 and l.startswith('    };')),default=0)
b=[m[0] for m in marks]
for k,v in agg.items():
    if k[0]=='lsq_kernels.cuh':
        i=bisect.bisect_right(b,k[1])-1
        name=marks[i][1][8:15] if i>=0 else 'pre'
        if emit0<=k[1]<=emit1: name='emit'
        ph[name][0]+=v[0]; ph[name][1]+=v[1]
    else:
        fn[k[0]][0]+=v[0]; fn[k[0]][1]+=v[1]
for n,v in list(ph.items())+list(fn.items()): print('%-28s %5.1f%% inst (%.0fM) %5.1f%% samp'%(n,100*v[0]/tot,v[0]/1e6,100*v[1]/ts))
N=int(sys.argv[2]) if len(sys.argv)>2 else 30
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:N]:
    print('%-22s %5d inst %5.1f%% samp %5.1f%% lanes %4.1f | %s'%(k[0],k[1],100*v[0]/tot,100*v[1]/max(ts,1),v[2]/max(v[0],1),src[k][:100]))
