import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from geoa3_b200 import ops
from geoa3_b200 import synth
from tools.time_kernels import timeit
b,n,k=250,1024,16
pc,nr,_=synth.make_batch(20,n)
ori=torch.from_numpy(np.tile(pc,(13,1,1))[:b].copy()).cuda(); nrm=torch.from_numpy(np.tile(nr,(13,1,1))[:b].copy()).cuda()
adv=ori+torch.from_numpy(synth.make_offsets(b,n)).cuda()
d1,js,d2,is_=ops.nn_pair(adv,ori); nbr=ops.knn(adv,adv,k+1,drop=1)[0]
ko=ops.kappa_loss_fwd(ori,normal=nrm,nbr=ops.knn(ori,ori,k+1,drop=1)[0])["kappa"]
out=ops.kappa_loss_fwd(adv,normal=nrm,jstar=js,nbr=nbr,d_a2o=d1,d_o2a=d2,kappa_ori=ko,want_nrm=True,want_cd=True,want_hd=True,want_curv=True)
g=torch.full((b,),1.0/b,device='cuda')
full=lambda: ops.loss_bwd(adv,ori=ori,nrm_adv=out["nrm"],kappa_adv=out["kappa"],kappa_ori=ko,jstar=js,istar=is_,nbr=nbr,hd_arg=out["hd_arg"],g_cd=g,g_hd=g,g_cu=g)
cd_only=lambda: ops.loss_bwd(adv,ori=ori,jstar=js,istar=is_,g_cd=g)
cd_one=lambda: ops.loss_bwd(adv,ori=ori,jstar=js,g_cd=g)
cu_only=lambda: ops.loss_bwd(adv,ori=ori,nrm_adv=out["nrm"],kappa_adv=out["kappa"],kappa_ori=ko,jstar=js,nbr=nbr,g_cu=g)
for name,f in (("full",full),("cd_two_sided",cd_only),("cd_one_sided",cd_one),("curv_only",cu_only)):
    print(name, timeit(f))
for bb in (74,148,296):
    a2,o2,n2=adv[:bb].contiguous(),ori[:bb].contiguous(),out["nrm"][:bb].contiguous()
    f=lambda: ops.loss_bwd(a2,ori=o2,nrm_adv=n2,kappa_adv=out["kappa"][:bb].contiguous(),kappa_ori=ko[:bb].contiguous(),jstar=js[:bb].contiguous(),istar=is_[:bb].contiguous(),nbr=nbr[:bb].contiguous(),hd_arg=out["hd_arg"][:bb].contiguous(),g_cd=g[:bb].contiguous(),g_hd=g[:bb].contiguous(),g_cu=g[:bb].contiguous()) if bb<=250 else None
    if bb<=250: print('b',bb, timeit(f))
import ctypes
from geoa3_b200 import _lib
lib=_lib.load()
try:
    buf=(ctypes.c_longlong*8)()
    full(); torch.cuda.synchronize(); ctypes.CDLL(_lib.SO_PATH).geoa3_debug_read(buf)
    print('phase cycles: load %d, csr1 %d, csr2 %d, accumulate %d | csr1(sorted): count %d scan %d fill %d sort %d'%tuple(buf[i] for i in range(8)))
except Exception as e: print('no debug', e)
