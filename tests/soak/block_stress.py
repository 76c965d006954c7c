#!/usr/bin/env python
"""Block-structure stress (test infrastructure): reads of 1-7 random blocks -- overlapping, nested, abutting, out of
order, nil, reversed, on other chromosomes/strands -- over a few genes with messy exon structures; the product's CPU
replay (tests/emul) against the oracle.  `solve` as third argument also compares the estimator.

    python tests/soak/block_stress.py 5000 800 [solve]   # first seed, number of runs (3 000 reads each)
"""
import sys, os, random, subprocess, tempfile, shutil
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,ROOT); sys.path.insert(0,ROOT+'/tests')
from oracle import runner
sys.path.insert(0,os.path.dirname(os.path.abspath(__file__)))
from live_fuzz import close
EMUL=ROOT+'/tests/emul/_build/emul'
seed0=int(sys.argv[1]); runs=int(sys.argv[2])
bad=0; hung=0; SOLVE=len(sys.argv)>3
for run in range(runs):
    rng=random.Random(seed0+run)
    d=tempfile.mkdtemp()
    # a few genes with messy exon structures on 2 chromosomes
    forms=[]; maps=[]
    for g in range(rng.randint(1,4)):
        chrom='chr%d'%rng.randint(1,2); base=rng.randint(100,3000)
        cuts=sorted(rng.sample(range(base,base+400),rng.randint(4,14)))
        nform=rng.randint(1,4)
        for f in range(nform):
            k=rng.randint(1,min(6,len(cuts)//2))
            pts=sorted(rng.sample(cuts,2*k))
            ex=[(pts[2*i],pts[2*i+1]) for i in range(k)]
            if rng.random()<0.3: rng.shuffle(ex)
            if rng.random()<0.2 and ex: ex.append((ex[0][0],ex[0][1]+rng.randint(0,5)))
            name='g%d|%d'%(g,f)
            forms.append('\t'.join([name,chrom,rng.choice('+-'),str(min(e[0] for e in ex)),str(max(e[1] for e in ex)),str(len(ex)),','.join(str(e[0]) for e in ex)+',',','.join(str(e[1]) for e in ex)+',']))
            maps.append('g%d\t%s'%(g,name))
    open(d+'/ev.interval','w').write('\n'.join(forms)+'\n'); open(d+'/ev.map','w').write('\n'.join(maps)+'\n')
    lines=['AlignmentBlocks']
    allcuts=[]
    for fl in forms:
        t=fl.split('\t'); st=[int(x) for x in t[6].split(',') if x]; en=[int(x) for x in t[7].split(',') if x]
        allcuts.append((t[1],t[2],sorted(set(st+en))))
    for i in range(3000):
        chrom,strand,cuts=rng.choice(allcuts)
        nb=rng.randint(1,7); blocks=[]
        for _ in range(nb):
            if rng.random()<0.7:
                a=rng.choice(cuts)+rng.randint(-3,3); b=a+rng.randint(1,40)
                if rng.random()<0.5: b=rng.choice(cuts)+rng.randint(-2,2)
            else:
                a=rng.randint(cuts[0]-10,cuts[-1]+10); b=a+rng.randint(0,30)
            if rng.random()<0.05: a,b=b,a
            a=max(a,0); b=max(b,0)
            blocks.append('%s:%s:%d:%d:1:%d'%(chrom if rng.random()<0.97 else 'chr9', strand if rng.random()<0.95 else '+', a+1, b, max(b-a,0)))
        if rng.random()<0.5: blocks.sort(key=lambda s:int(s.split(':')[2]))
        lines.append(','.join(blocks))
    open(d+'/r.mrf','w').write('\n'.join(lines)+'\n')
    R=rng.choice([25,36,50,75])
    a=['0','p',d+'/o_','LH_GENE_TXT',d+'/ev.interval','UCSC_GENE2ISOFORM',d+'/ev.map','0','100','MRF_SINGLE',rng.choice(['SHORT_READ','MEDIUM_READ']),str(R),d+'/r.mrf']
    rc1,o1,e1=runner.run_oracle('count',a,timeout=60)
    p=subprocess.run([EMUL,'count']+a,capture_output=True,text=True)
    if rc1 in (0,) and (p.returncode!=rc1 or p.stdout!=o1):
        bad+=1; keep=os.path.join(tempfile.gettempdir(),'lsq_stress_fail_%d'%(seed0+run)); shutil.copytree(d,keep,dirs_exist_ok=True); print('MISMATCH run',seed0+run,'oracle rc',rc1,'emul rc',p.returncode,'kept',keep,flush=True)
    elif rc1!=0 and run<3: print('oracle rc',rc1,e1[-100:])
    elif rc1==0 and SOLVE:
        a2=a+['1000000.0']
        try:
            rc3,o3,_=runner.run_oracle('solve',a2,timeout=5)
            env=dict(os.environ); env['LSQ_PRECISION']='17'
            p2=subprocess.run([EMUL,'solve']+a2,env=env,capture_output=True,text=True,timeout=30)
            if rc3==0 and (p2.returncode!=0 or not close(o3,p2.stdout,True)):
                bad+=1; keep=os.path.join(tempfile.gettempdir(),'lsq_stress_fail_s%d'%(seed0+run)); shutil.copytree(d,keep,dirs_exist_ok=True); print('SOLVE MISMATCH run',seed0+run,'kept',keep,flush=True)
        except subprocess.TimeoutExpired: hung+=1
    shutil.rmtree(d,ignore_errors=True)
print('runs',runs,'mismatches',bad,'hung',hung)
