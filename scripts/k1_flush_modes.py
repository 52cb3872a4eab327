import ctypes as C, os, sys
ROOT='/root/repo'
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200"), os.path.join(ROOT,'scripts')): sys.path.insert(0, p)
import numpy as np, torch
from ckb200 import lib as L
import bench_sweep as BS
lib=L.raw()
full=BS.make_positions(1<<22)
stream=torch.cuda.current_stream().cuda_stream
wflush=torch.empty(256<<20,dtype=torch.uint8,device='cuda')
rflush=torch.ones(512<<20,dtype=torch.uint8,device='cuda')
for lg in (20,22):
    n=1<<lg
    pos=full[:n]
    d_pos=torch.from_numpy(pos.view(np.uint32).reshape(n,4).copy()).cuda()
    d_counts=torch.zeros(n,dtype=torch.int32,device='cuda')
    d_masks=torch.zeros((n,8),dtype=torch.int32,device='cuda'); d_status=torch.zeros(n,dtype=torch.uint8,device='cuda'); d_p5=torch.zeros(n,dtype=torch.uint8,device='cuda')
    d_off=torch.zeros(n+1,dtype=torch.int32,device='cuda')
    L.check(lib.ck_movegen_csr_device(C.c_void_p(d_pos.data_ptr()), n, None, 0, C.c_void_p(d_off.data_ptr()), None,None,None, C.c_void_p(stream)))
    torch.cuda.synchronize(); total=int(d_off[-1].item())
    d_packed=torch.zeros((total,4),dtype=torch.int32,device='cuda')
    def launch():
        L.check(lib.ck_movegen_csr_device(C.c_void_p(d_pos.data_ptr()), n, C.c_void_p(d_packed.data_ptr()), total, C.c_void_p(d_off.data_ptr()), C.c_void_p(d_masks.data_ptr()), C.c_void_p(d_status.data_ptr()), C.c_void_p(d_p5.data_ptr()), C.c_void_p(stream)))
    for mode in ('none','write','read'):
        ts=[]
        for it in range(12):
            if mode=='write': wflush.fill_(it)
            elif mode=='read': s=rflush.sum()
            e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            e0.record(); launch(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms=float(np.median(ts[2:])); byt=n*(16+4+32+2)+total*16
        print('n=2^%d flush=%s: %.4f ms  %.1f GB/s  children %.2f/pos'%(lg,mode,ms,byt/ms/1e6,total/n),flush=True)
