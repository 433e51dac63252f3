#!/usr/bin/env python
"""Offline accuracy study behind `bin_mass` (hesic_b200/csrc/elementwise.cu): the probability mass of a quantisation bin,
Phi((0.5 - d) / s) - Phi((-0.5 - d) / s), evaluated in emulated fp32 three ways -- the reference's way (torch.erfc differences),
a cheaper erfc (Chebyshev fit, 1.2e-7 relative), and the hybrid the kernel uses (series in 1 / s for narrow bins, the cheap erfc
otherwise) -- against the exact value (float64 scipy).  Relative errors are taken with the tests' floor of 1e-9.  CPU only.

    python tools/erfc_eval.py"""
import numpy as np, torch, math
f32=np.float32
def erfcc(x):  # NR erfcc in float32 ops
    x=x.astype(f32); z=np.abs(x)
    t=(f32(1)/(f32(1)+f32(0.5)*z)).astype(f32)
    c=[0.17087277,-0.82215223,1.48851587,-1.13520398,0.27886807,-0.18628806,0.09678418,0.37409196,1.00002368]
    p=f32(c[0])*np.ones_like(t)
    for k in c[1:]:
        p=(p*t+f32(k)).astype(f32)
    e=(-z*z-f32(1.26551223)+t*p).astype(f32)
    r=(t*np.exp2((e*f32(1.4426950408889634)).astype(f32)).astype(f32)).astype(f32)
    return np.where(x>=0,r,f32(2)-r).astype(f32)
rng=np.random.default_rng(0)
N=4_000_000
sig=np.exp(rng.uniform(np.log(0.11),np.log(300),N)).astype(f32)
d=(np.abs(rng.standard_normal(N))*sig*2.5).astype(f32)   # |q-mu|
d=np.where(rng.random(N)<0.3, np.floor(d), d).astype(f32)
inv=(f32(1)/sig).astype(f32)
u=((f32(0.5)-d)/sig).astype(f32); l=((f32(-0.5)-d)/sig).astype(f32)
k=f32(0.70710678118654752440)
# reference path (torch fp32, like the oracle)
tu=torch.from_numpy(u); tl=torch.from_numpy(l)
ref=(0.5*torch.erfc(-k*tu)-0.5*torch.erfc(-k*tl)).numpy()
# exact (float64)
from scipy.special import erfc as erfc64
ex=0.5*erfc64(-(0.5-d.astype(np.float64))/sig.astype(np.float64)/math.sqrt(2))-0.5*erfc64(-(-0.5-d.astype(np.float64))/sig.astype(np.float64)/math.sqrt(2))
# fast
fast=(f32(0.5)*erfcc((-k*u).astype(f32))-f32(0.5)*erfcc((-k*l).astype(f32))).astype(f32)
u2=((f32(0.5)-d)*inv).astype(f32); l2=((f32(-0.5)-d)*inv).astype(f32)
fast2=(f32(0.5)*erfcc((-k*u2).astype(f32))-f32(0.5)*erfcc((-k*l2).astype(f32))).astype(f32)
def rel(a,b): return np.abs(a.astype(np.float64)-b.astype(np.float64))/np.maximum(np.abs(b.astype(np.float64)),1e-9)
for name,a in (("torch vs exact",ref),("fast vs exact",fast),("fast(rcp) vs exact",fast2)):
    r=rel(a,ex); print(f"{name:22s} max {r.max():.3e}  99.99% {np.quantile(r,0.9999):.3e}  mean {r.mean():.3e}")
r=rel(fast2,ref); print(f"fast(rcp) vs torch     max {r.max():.3e}  99.99% {np.quantile(r,0.9999):.3e}  frac>1e-4 {np.mean(r>1e-4):.2e}")
i=np.argmax(r); print("worst:", sig[i], d[i], ref[i], fast2[i], ex[i])
for smax in (2,8,32,128):
    m=sig<smax; rr=rel(fast2[m],ref[m]); print(f" sigma<{smax}: max {rr.max():.3e} frac>1e-4 {np.mean(rr>1e-4):.2e};   torch vs exact max {rel(ref[m],ex[m]).max():.3e}")

def erfc_pos(z):
    z=z.astype(f32)
    t=(f32(1)/(f32(1)+f32(0.5)*z)).astype(f32)
    c=[0.17087277,-0.82215223,1.48851587,-1.13520398,0.27886807,-0.18628806,0.09678418,0.37409196,1.00002368]
    p=f32(c[0])*np.ones_like(t)
    for kk in c[1:]:
        p=(p*t+f32(kk)).astype(f32)
    e=(-z*z+(t*p-f32(1.26551223))).astype(f32)
    return (t*np.exp2((e*f32(1.4426950408889634)).astype(f32)).astype(f32)).astype(f32)
def hybrid(d,inv,thr=0.125):
    h=inv
    m=(d*h).astype(f32); m2=(m*m).astype(f32)
    phi=(f32(0.3989422804014327)*np.exp2((m2*f32(-0.5*1.4426950408889634)).astype(f32)).astype(f32)).astype(f32)
    h2=(h*h).astype(f32)
    c=(f32(1)+h2*((m2-f32(1))*f32(1/24)+h2*((m2*(m2-f32(6))+f32(3))*f32(1/1920)))).astype(f32)
    ser=(h*phi*c).astype(f32)
    a=((d-f32(0.5))*h*k).astype(f32); b=((d+f32(0.5))*h*k).astype(f32)
    eb=erfc_pos(b); ea=np.where(a>=0,erfc_pos(np.abs(a)),f32(2)-erfc_pos(np.abs(a))).astype(f32)
    er=(f32(0.5)*(ea-eb)).astype(f32)
    return np.where(h<=f32(thr),ser,er).astype(f32)
for thr in (0.125,0.25,0.5):
    hy=hybrid(d,inv,thr)
    r1=rel(hy,ex); r2=rel(hy,ref)
    print(f"thr {thr}: hybrid vs exact max {r1.max():.3e} 99.99% {np.quantile(r1,0.9999):.3e} | vs torch max {r2.max():.3e} 99.99% {np.quantile(r2,0.9999):.3e} frac>1e-4 {np.mean(r2>1e-4):.2e}")
    for smax in (2,8,32,128):
        m_=sig<smax; print(f"   sigma<{smax}: vs exact {r1[m_].max():.3e}  vs torch {r2[m_].max():.3e}")
