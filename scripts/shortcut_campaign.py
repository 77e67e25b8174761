"""Large-scale CPU check of the Seidel-shortcut model (oracle/shortcut_model.c) against the sequential restatement:
random 7-DOF / 200-gridpoint paths in four flavours (benchmark-like, velocity-limited, randomly scaled waypoints and
limits, non-uniform grids), 4096 paths per job, every 2-variable LP compared bit for bit.
usage: python scripts/shortcut_campaign.py <processes> <jobs>      (10000 jobs = 1.63e10 LPs, about 25 min on 8 cores)"""
import sys, time, numpy as np, multiprocessing as mp
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
def work(args):
    seed, B, kind = args
    from oracle import oracle as orc
    from problems import make_batch_fast
    rng=np.random.RandomState(seed)
    with orc.shortcut_model() as sm:
        grid=np.linspace(0,1,200)
        ss,way,vlim,alim=make_batch_fast(B, seed)
        if kind==1: vlim=vlim*0.03           # velocity-limited
        if kind==2: way=way*rng.uniform(0.01,3,size=(B,1,1)); alim=alim*rng.uniform(0.05,5,size=(B,1,1))
        if kind==3: grid=np.sort(np.r_[0,rng.rand(198),1.0])   # non-uniform grid
        c=np.stack([orc.cubic_spline_fit(ss,way[b]) for b in range(B)])
        r=orc.solve_velacc_batch(c,np.tile(ss,(B,1)),grid,vlim,alim,nthreads=1)
        st=sm.stats()
    return st['lps'], st['mismatches'], st['a_used'], st['a_declined'], st['b_used'], st['b_declined'], int((r['status']!=0).sum())
if __name__=='__main__':
    nproc=int(sys.argv[1]); njobs=int(sys.argv[2]); B=4096
    t=time.time(); tot=np.zeros(7,dtype=np.int64)
    with mp.Pool(nproc) as pool:
        for i,res in enumerate(pool.imap_unordered(work, [(5000000+j, B, j%4) for j in range(njobs)])):
            tot+=np.array(res)
            if (i+1)%500==0: print(i+1,'jobs', tot, '%.0fs'%(time.time()-t), flush=True)
    print('TOTAL lps mismatches a_used a_decl b_used b_decl failed_paths:', tot, '%.0fs'%(time.time()-t))
