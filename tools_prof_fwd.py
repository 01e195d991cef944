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
t0=a[:,0].min()
tot=a[:,7]
d=np.diff(a[:,:7],axis=1)
names=['geom','stage+tab','own items','extras+combine','(round0 end)->loop end','res write+sync']
print('CTA lifetime mean', (a[:,6]-a[:,0]).mean(), 'max', (a[:,6]-a[:,0]).max(), ' kernel span', (a[:,6].max()-t0))
for k,n in enumerate(names): print(f'{n:28s} mean {d[:,k].mean():9.0f}  p50 {np.median(d[:,k]):9.0f} max {d[:,k].max():9.0f}')
sel=(tot>500)&(tot<900)
print('typical tiles (500-900 pts):', sel.sum())
for k,n in enumerate(names): print(f'   {n:28s} mean {d[sel,k].mean():9.0f}')
# start time distribution
st=np.sort(a[:,0]-t0)
print('start times: p10 %d p50 %d p90 %d max %d' % (st[409], st[2048], st[3686], st[-1]))
heavy=np.argsort(-tot)[:8]
for h in heavy: print('heavy tile', h, 'pts', tot[h], 'start', a[h,0]-t0, 'life', a[h,6]-a[h,0])
