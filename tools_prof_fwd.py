import ctypes, sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
import bench
from rcbevdet_b200 import _lib, rig, bev_pool as bp
from rcbevdet_b200.prepare import prepare_async
dev=torch.device('cuda',0)
coor,depth,feat,og = bench.make_inputs(torch, rig, dev, 8, 1)
lo,iv,sz = rig.grid_tensors(rig.R50_GRID)
import rcbevdet_b200 as rcb
for it in range(3):
    bev = rcb.voxel_pooling_v2(coor, depth, feat, lo, iv, sz)
torch.cuda.synchronize()
lib=_lib.lib()
buf=(ctypes.c_longlong*(8192*8))()
lib.rcb_debug_fwd_prof.argtypes=[ctypes.c_void_p, ctypes.c_int]
print('rc', lib.rcb_debug_fwd_prof(buf, 8192*8))
a=np.array(buf[:4096*8]).reshape(4096,8)
tot=a[:,7]
ne=tot>0
life=(a[:,4]-a[:,0])
print('nonempty tiles', ne.sum(), 'life mean', life[ne].mean(), 'max', life[ne].max(), 'empty life', life[~ne].mean() if (~ne).any() else 0)
names=['geom','stage+table (round0)','items (round0)','combine+later rounds']
d=np.diff(a[:,:5],axis=1)
for lo,hi in [(1,300),(300,900),(900,1500),(1500,3000),(3000,10000)]:
    sel=(tot>=lo)&(tot<hi)
    if sel.sum()==0: continue
    print(f'tiles with {lo}-{hi} pts: {sel.sum()}  life {life[sel].mean():.0f}  ' + '  '.join(f'{n} {d[sel,k].mean():.0f}' for k,n in enumerate(names)))
print('sum of lifetimes / (148*3 slots) =', life.sum()/(148*3), 'cycles')
