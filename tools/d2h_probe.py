import torch, time, numpy as np
n=15120000
d=torch.empty(n,dtype=torch.uint8,device='cuda')
pin=torch.empty(n,dtype=torch.uint8).pin_memory()
pag=torch.empty(n,dtype=torch.uint8)
pag2=np.empty(n,np.uint8)
def t(f,reps=20):
    f(); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter()-t0)/reps*1e3
print('d2h pinned ms', t(lambda: pin.copy_(d,non_blocking=True)))
print('d2h pageable ms', t(lambda: pag.copy_(d)))
pn=pin.numpy()
print('memcpy pinned->pageable 1 thread ms', t(lambda: np.copyto(pag2,pn)))
import threading
def par(k):
    step=(n+k-1)//k
    ths=[threading.Thread(target=lambda a=a: np.copyto(pag2[a:a+step],pn[a:a+step])) for a in range(0,n,step)]
    [x.start() for x in ths]; [x.join() for x in ths]
for k in (2,4,8): print('memcpy',k,'threads ms', t(lambda: par(k)))
fresh=lambda: np.copyto(np.empty(n,np.uint8),pn)
print('memcpy into fresh array ms', t(fresh))
