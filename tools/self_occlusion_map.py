"""CPU study (oracle traversal, numpy float32 statement of the light sampling): how often the light's own triangles occlude an otherwise
visible light sample under the reference's shadow test (t > 1e-5 and t_to_light - t > 1e-5, Render.cuh:19-27), per 8x8 block of the
240x180 cornell-box frame. Output quoted in profiles/r02_converged.md."""
import sys, os, tempfile, math
sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import orc
from tools import scene_fixture as sf
import bench
cfg=bench.Workload("c1", use_product_loader=False)
W,H=240,180
O=orc.Scene().add_obj(cfg.obj,cfg.tmp); O.build_new_bvh(2)
M=orc.inverse_view_matrix(cfg.eye,cfg.lookat,cfg.up)
rays=orc.primary_rays(cfg.eye,M,float(cfg.fovy_rad),W,H)
t,face=O.trace(rays,which=0)
tri=O.tris(); lights=O.lights()
lf=lights[0][0]
is_light=np.zeros(O.n_tris,bool); is_light[lf]=True
f32=np.float32
hitp=(rays[:,0:3]+t[:,None]*rays[:,4:7]).astype(f32)
valid=(face>=0)&(~is_light[np.maximum(face,0)])
rng=np.random.default_rng(1)
S=32
self_occ=np.zeros(W*H); reach=np.zeros(W*H)
V=tri["verts"][lf].reshape(-1,3,3)
ln=tri["normal"][lf]
nrm=tri["normal"][np.maximum(face,0)]
def norm32(v):
    n=(v[:,0]*v[:,0]); n=(v[:,1]*v[:,1]+n).astype(f32); n=(v[:,2]*v[:,2]+n).astype(f32)
    return (v/np.sqrt(n).astype(f32)[:,None]).astype(f32)
for s in range(S):
    k=rng.integers(0,len(lf),W*H)
    a=rng.random(W*H).astype(f32); b=(rng.random(W*H).astype(f32)*(f32(1)-a)).astype(f32); g=((f32(1)-a)-b).astype(f32)
    lp=((a[:,None]*V[k,0]+b[:,None]*V[k,1]).astype(f32)+g[:,None]*V[k,2]).astype(f32)
    dist=(lp-hitp).astype(f32)
    d=norm32(dist)
    ttl=(dist[:,0]/d[:,0]).astype(f32)
    cos1=(d*nrm).sum(1); cos2=-(d*ln[k]).sum(1)
    ok=valid&(cos1>0)&(cos2>0)&np.isfinite(ttl)
    sh=np.zeros((W*H,8),f32); sh[:,0:3]=hitp; sh[:,3]=ttl; sh[:,4:7]=norm32(d)
    sh[~ok,4:7]=[0,0,1]; sh[~ok,3]=1e-3
    tb,fb=O.trace(sh,which=0,mode=1)
    blocked=ok&(fb>=0)
    selfb=blocked&is_light[np.maximum(fb,0)]
    # would it be visible if the light's own triangles are ignored? approximate: count self-blocked among all ok samples not blocked by non-light geometry
    reach+=ok&((fb<0)|selfb); self_occ+=selfb
rate=np.where(reach>0,self_occ/np.maximum(reach,1),np.nan).reshape(H,W)
print("overall self-occlusion rate of otherwise visible light samples: %.4f"%(np.nansum(self_occ)/np.nansum(reach)))
def box(img,k=8):
    h,w=img.shape; return np.nanmean(img[:h-h%k,:w-w%k].reshape(h//k,k,w//k,k),axis=(1,3))
rb=box(rate)

for r in rb[::2]: print(" ".join("%3d"%(0 if np.isnan(x) else round(100*x)) for x in r))
