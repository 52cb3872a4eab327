import os, sys
ROOT='/root/repo'
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")): sys.path.insert(0, p)
import numpy as np
from ckb200 import lib as L, net as N
n=int(sys.argv[1]); out=sys.argv[2]
rng=np.random.RandomState(0)
pos=np.zeros(n,dtype=L.POS_DTYPE); pos["p1"],pos["p2"]=0x00000FFF,0xFFF00000
for _ in range(14):
    o=L.movegen(pos); pick=(rng.rand(n)*np.maximum(o["counts"],1)).astype(np.int64)
    nxt=o["children"][np.arange(n),pick]; alive=(o["status"]==0)&(o["counts"]>0); pos=np.where(alive,nxt,pos)
o=L.movegen(pos,want_children=False)
leaves=np.zeros(n,dtype=L.LEAF_DTYPE); leaves["p1"],leaves["p2"],leaves["k"]=pos["p1"],pos["p2"],pos["k"]; leaves["info"]=pos["meta"]&1; leaves["mask"]=o["masks"]
net=L.Net(0,"tc"); net.set_weights(N.random_init_blob(0))
res=[net.forward(leaves) for _ in range(3)]
for i in (1,2):
    assert res[i][0].tobytes()==res[0][0].tobytes() and res[i][1].tobytes()==res[0][1].tobytes(), "not reproducible run to run"
np.savez(out,pol=res[0][0],val=res[0][1])
print("saved",out)
