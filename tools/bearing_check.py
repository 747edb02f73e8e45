"""Development check of bearing_atan2 (csrc/mh_costs.cuh) in emulated float32: octant reduction, an approximate division
(modelled as the correctly rounded reciprocal moved by one ulp at random), the degree-7 polynomial with fused multiply-adds.
Prints the largest absolute error against a float64 atan2 over 4e6 points (incl. axis-aligned, diagonal and tiny inputs)."""
import numpy as np
c = np.float32([0.9999993443489075, -0.33329859375953674, 0.19946561753749847, -0.1390860676765442, 0.09642140567302704, -0.0559115894138813, 0.021862473338842392, -0.004054440185427666])
def fma(a,b,cc): return (a.astype(np.float64)*b.astype(np.float64)+cc.astype(np.float64)).astype(np.float32)
def bearing(y,x):
    ax=np.abs(x); ay=np.abs(y)
    mx=np.maximum(ax,ay); mn=np.minimum(ax,ay)
    # rcp.approx: 1 ulp error model: exact reciprocal rounded, then perturb by +-1ulp randomly
    r=(1.0/mx.astype(np.float64)).astype(np.float32)
    r=np.nextafter(r, np.where(np.random.rand(*r.shape)<0.5, np.float32(0), np.float32(np.inf))).astype(np.float32)
    t=(mn*r).astype(np.float32)
    t=np.where(mx==0, np.float32(0), t)
    s=(t*t).astype(np.float32)
    p=np.full_like(s, c[7])
    for k in range(6,-1,-1): p=fma(p,s,np.full_like(s,c[k]))
    a=(t*p).astype(np.float32)
    a=np.where(ay>ax, np.float32(np.pi/2)-a, a).astype(np.float32)
    a=np.where(x<0, np.float32(np.pi)-a, a).astype(np.float32)
    return np.where(y<0, -a, a).astype(np.float32)
rng=np.random.default_rng(1)
N=4_000_000
x=rng.uniform(-20,20,N).astype(np.float32); y=rng.uniform(-20,20,N).astype(np.float32)
# also small / axis-aligned cases
x[:1000]=0; y[1000:2000]=0; x[2000:3000]=y[2000:3000]; x[3000:4000]*=1e-6
got=bearing(y,x); ref=np.arctan2(y.astype(np.float64),x.astype(np.float64))
err=np.abs(got.astype(np.float64)-ref)
ulp=np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
print("max abs err", err.max(), "max ulp err", (err/np.maximum(ulp,1e-45)).max(), "p99.99 ulp", np.quantile(err/np.maximum(ulp,1e-45),0.9999))
# libm atan2f-like: correctly rounded float of ref
f32=ref.astype(np.float32)
print("vs correctly rounded: max diff in ulps", (np.abs(got.astype(np.float64)-f32.astype(np.float64))/np.maximum(ulp,1e-45)).max())
i=np.argmax(err/np.maximum(ulp,1e-45)); print(x[i],y[i],got[i],ref[i])
