import ctypes, sys, numpy as np, torch
# Debug helper: per-CTA phase timeline of k_pool_fwd_tile.  Needs a build with -DRCB_PROFILE_PHASES
# (tools/build_variant.sh prof "-DRCB_PROFILE_PHASES", then RCB_LIB_PATH=rcbevdet_b200/lib/variants/lib_prof.so);
# run from the repo root on the GPU box.
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import bench
from rcbevdet_b200 import _lib, rig
import rcbevdet_b200 as rcb
dev=torch.device('cuda',0)
coor,depth,feat,og = bench.make_inputs(torch, rig, dev, 8, 1)
lo,iv,sz = rig.grid_tensors(rig.R50_GRID)
lib=_lib.lib()
lib.rcb_debug_fwd_prof.argtypes=[ctypes.c_void_p, ctypes.c_int]
for it in range(3):
    bev = rcb.voxel_pooling_v2(coor, depth, feat, lo, iv, sz)
torch.cuda.synchronize()
lib.rcb_debug_fwd_prof_reset()
bev = rcb.voxel_pooling_v2(coor, depth, feat, lo, iv, sz)
torch.cuda.synchronize()
N=16384
buf=(ctypes.c_longlong*(N*8))()
print('rc', lib.rcb_debug_fwd_prof(buf, N*8))
a=np.array(buf[:]).reshape(N,8)
a=a[a[:,0]>0]
t0=a[:,0].min(); tend=a[:,5].max()
print('CTAs that ran', len(a), 'kernel span us', (tend-t0)/1e3)
life=(a[:,5]-a[:,0])/1e3
tot=a[:,7]; sm=a[:,6]
print('life us: mean %.2f p50 %.2f p90 %.2f max %.2f' % (life.mean(), np.median(life), np.percentile(life,90), life.max()))
for lo_,hi_ in [(0,1),(1,300),(300,900),(900,1500),(1500,3000),(3000,10000)]:
    sel=(tot>=lo_)&(tot<hi_)
    if sel.sum(): print(f'  tiles {lo_}-{hi_} pts: n={sel.sum()} life mean {life[sel].mean():.2f} us; phases us geom {((a[sel,1]-a[sel,0])/1e3).mean():.2f} stage {((a[sel,2]-a[sel,1])/1e3).mean():.2f} items {((a[sel,3]-a[sel,2])/1e3).mean():.2f} rest {((a[sel,4]-a[sel,3])/1e3).mean():.2f} write {((a[sel,5]-a[sel,4])/1e3).mean():.2f}')
# per-SM busy analysis
span=(tend-t0)
conc=[]
for s_ in np.unique(sm):
    sel=sm==s_
    ev=np.concatenate([np.stack([a[sel,0],np.ones(sel.sum())],1), np.stack([a[sel,5],-np.ones(sel.sum())],1)])
    ev=ev[np.argsort(ev[:,0])]
    c=np.cumsum(ev[:,1]); dt=np.diff(ev[:,0]); 
    avg=(c[:-1]*dt).sum()/span
    last=a[sel,5].max()-t0
    conc.append((avg,last/1e3,sel.sum()))
conc=np.array(conc)
print('per-SM: avg concurrent CTAs mean %.2f min %.2f; last-finish us mean %.1f min %.1f max %.1f; CTAs/SM mean %.1f' % (conc[:,0].mean(), conc[:,0].min(), conc[:,1].mean(), conc[:,1].min(), conc[:,1].max(), conc[:,2].mean()))
st=np.sort((a[:,0]-t0)/1e3)
print('CTA start times us: p10 %.1f p50 %.1f p90 %.1f max %.1f' % (st[len(st)//10], st[len(st)//2], st[len(st)*9//10], st[-1]))
hv=np.argsort(-tot)[:6]
for h in hv: print('  heavy: pts', tot[h], 'start %.1f life %.1f' % ((a[h,0]-t0)/1e3, life[h]))
