import sys, numpy as np
sys.path.insert(0,'/root/repo'); 
from __graft_entry__ import load_pkg
fv=load_pkg(); ctx=fv.Context(0)
rng=np.random.RandomState(0)
for K,B in ((3965,128),(3965,32),(3965,8),(3965,128)):
    s=np.float32(-rng.uniform(0,30,K))
    hv,hs=ctx.heap_replay(s,B)
# sorted descending: no replacements after fill
s=np.sort(np.float32(-rng.uniform(0,30,3965)))[::-1].copy(); ctx.heap_replay(s,128)
# ascending: every element replaces
s=np.sort(np.float32(-rng.uniform(0,30,3965))).copy(); ctx.heap_replay(s,128)
ctx.sync()
