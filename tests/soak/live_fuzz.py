#!/usr/bin/env python
"""Live differential soak (test infrastructure, dev container only): the product's CPU replay (tests/emul: host front
end + table layout + device functions) against the REFERENCE'S OWN BINARIES (oracle/_ref) on fresh fuzz seeds.

    python tests/soak/live_fuzz.py 30000 42000        # seeds [lo, hi): count and solve for each, 17 digits

A reference run that does not finish in 5 s is reported as hung (its EM loop can cycle for ever on events whose
log-likelihood is rounding noise).  Round 1 found the filter_read_t spill bug this way.
"""
import sys, os, tempfile, shutil, subprocess, math, time
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,ROOT); sys.path.insert(0,ROOT+'/tests')
import fuzzgen
from oracle import runner
EMUL=ROOT+'/tests/emul/_build/emul'
def close(a,b,solve):
    la,lb=a.splitlines(),b.splitlines()
    if len(la)!=len(lb): return False
    for x,y in zip(la,lb):
        fx,fy=x.split('\t'),y.split('\t')
        if len(fx)!=len(fy): return False
        for u,v in zip(fx,fy):
            if u==v: continue
            if not solve: return False
            try: du,dv=float(u),float(v)
            except ValueError: return False
            if math.isnan(du) and math.isnan(dv): continue
            if abs(du-dv)>1e-9*max(abs(du),abs(dv)) and abs(du-dv)>1e-12: return False
    return True
def main():
    lo,hi=int(sys.argv[1]),int(sys.argv[2])
    bad=0; checked=0; t0=time.time(); hung=[]
    for seed in range(lo,hi):
        case=fuzzgen.make_case(seed)
        d=tempfile.mkdtemp()
        try:
            ca,sa=fuzzgen.write_case(case,d)
            for prog,args in (('count',ca),('solve',sa)):
                try: rc1,o1,_=runner.run_ref(prog,args,timeout=5)
                except subprocess.TimeoutExpired: hung.append(seed); continue
                if rc1!=0: continue
                env=dict(os.environ); env['LSQ_PRECISION']='17'
                p=subprocess.run([EMUL,prog]+[str(a) for a in args],env=env,capture_output=True,text=True)
                checked+=1
                if p.returncode!=0 or not close(o1,p.stdout,prog=='solve'):
                    bad+=1; print('MISMATCH seed',seed,prog,'rc',p.returncode,flush=True)
        finally:
            shutil.rmtree(d,ignore_errors=True)
    print('seeds',lo,hi,'checked',checked,'mismatches',bad,'reference hung on',hung,'%.0fs'%(time.time()-t0),flush=True)


if __name__ == '__main__':
    main()
