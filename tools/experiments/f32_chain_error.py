"""Error of a float32 NeRF chain (rotation and position in float32, coordinates relative to the anchor) against float64, as a
function of the extent of the chain: the measurement behind the float32 first pass of backmap_fwd6_kernel and its 16 nm
extent limit (python tools/experiments/f32_chain_error.py; numpy only, ~2 minutes)."""
import numpy as np
f32=np.float32
def chain32(dih, ang, L):
    B, n = dih.shape
    R = np.tile(np.eye(3, dtype=f32), (B,1,1)); p = np.zeros((B,3), f32); out = np.zeros((B,n,3), f32)
    for k in range(n):
        phi = dih[:,k].astype(np.float64); th = ang[:,k].astype(np.float64)
        cw, sw = np.cos(phi).astype(f32), np.sin(phi).astype(f32)
        cg, sg = (-np.cos(th)).astype(f32), np.sin(th).astype(f32)
        c1 = (R[:,:,1]*cw[:,None]).astype(f32) + (R[:,:,2]*sw[:,None]).astype(f32)   # unfused worst case
        c2 = (R[:,:,2]*cw[:,None]).astype(f32) - (R[:,:,1]*sw[:,None]).astype(f32)
        R[:,:,1], R[:,:,2] = c1, c2
        c0 = (R[:,:,0]*cg[:,None]).astype(f32) + (R[:,:,1]*sg[:,None]).astype(f32)
        c1 = (R[:,:,1]*cg[:,None]).astype(f32) - (R[:,:,0]*sg[:,None]).astype(f32)
        R[:,:,0], R[:,:,1] = c0, c1
        p = (p + (f32(L[k])*R[:,:,0]).astype(f32)).astype(f32)
        out[:,k] = p
    return out
def chain64(dih, ang, L):
    B, n = dih.shape
    R = np.tile(np.eye(3), (B,1,1)); p = np.zeros((B,3)); out = np.zeros((B,n,3))
    for k in range(n):
        phi = dih[:,k].astype(np.float64); th = ang[:,k].astype(np.float64)
        cw, sw, cg, sg = np.cos(phi), np.sin(phi), -np.cos(th), np.sin(th)
        c1 = R[:,:,1]*cw[:,None] + R[:,:,2]*sw[:,None]; c2 = R[:,:,2]*cw[:,None] - R[:,:,1]*sw[:,None]
        R[:,:,1], R[:,:,2] = c1, c2
        c0 = R[:,:,0]*cg[:,None] + R[:,:,1]*sg[:,None]; c1 = R[:,:,1]*cg[:,None] - R[:,:,0]*sg[:,None]
        R[:,:,0], R[:,:,1] = c0, c1
        p = p + L[k]*R[:,:,0]; out[:,k] = p
    return out
rng = np.random.default_rng(1)
B, n = 4096, 750
L = rng.uniform(0.13,0.15,n).astype(f32).astype(np.float64)
ang = rng.uniform(1.9,2.2,(B,n)).astype(f32)
for name, dih in (("random", rng.uniform(-np.pi,np.pi,(B,n))), ("narrow", rng.normal(2.5,0.3,(B,n))), ("helix", np.tile(np.array([-1.0,-0.8,np.pi]),(B,250))+rng.normal(0,0.05,(B,n)))):
    dih = dih.astype(f32)
    ref = chain64(dih, ang, L); t = chain32(dih, ang, L)
    err = np.linalg.norm(t-ref,axis=2).max(1); ext = np.linalg.norm(ref,axis=2).max(1)
    print(name, "max err %.2e  p99.9 %.2e  extent max %.1f median %.1f  max err/extent %.2e" % (err.max(), np.quantile(err,0.999), ext.max(), np.median(ext), (err/ext).max()))
    for thr in (12,16,20):
        m = ext<thr
        if m.any(): print("   extent<%d: %5.1f%% of frames, max err %.2e" % (thr, 100*m.mean(), err[m].max()))
