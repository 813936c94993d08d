"""How exactly does fp32 evaluate the SSIM term?  200 000 synthetic 3x3 windows: the reference's operation order (sequential
window sums, /9, sigma = E[x^2] - mu^2, multiview_photometric_loss.py:35-51) and the grouped loss program's formula (separable
sums, quotient on sums scaled by 81^2, csrc/loss_group_kernel.cuh) against float64.  Both carry the SAME cancellation noise
(median 6e-6, 99 % 9e-5 absolute on a value in [0,1]): two candidates of the per-pixel minimum that are closer than that
resolve arbitrarily in EITHER implementation, which is why the gradient parity tests allow a few 3x3 neighbourhoods of
kink pixels (tests/test_loss_gpu.py assert_field_close).  CPU only: python tools/ssim_noise_study.py"""
import numpy as np
rng=np.random.default_rng(0)
n=200000
# smooth-ish windows: base level + small variation
base=rng.uniform(0.1,0.9,(n,1)).astype(np.float32)
amp=rng.choice([0.002,0.01,0.05],(n,1)).astype(np.float32)
x=(base+amp*rng.standard_normal((n,9))).astype(np.float32).clip(0,1)
y=(base+amp*rng.standard_normal((n,9))+0.01).astype(np.float32).clip(0,1)
C1,C2=np.float32(1e-4),np.float32(9e-4)
def exact(x,y):
    x=x.astype(np.float64); y=y.astype(np.float64)
    mx=x.mean(1); my=y.mean(1); sx=(x*x).mean(1)-mx*mx; sy=(y*y).mean(1)-my*my; sxy=(x*y).mean(1)-mx*my
    s=((2*mx*my+1e-4)*(2*sxy+9e-4))/((mx*mx+my*my+1e-4)*(sx+sy+9e-4))
    return np.clip((1-s)/2,0,1)
def ref32(x,y):
    f=np.float32
    def seqsum(a):
        s=a[:,0].copy()
        for i in range(1,9): s=(s+a[:,i]).astype(f)
        return s
    mx=(seqsum(x)/f(9)).astype(f); my=(seqsum(y)/f(9)).astype(f)
    sx=((seqsum((x*x).astype(f))/f(9)).astype(f)-(mx*mx).astype(f)).astype(f)
    sy=((seqsum((y*y).astype(f))/f(9)).astype(f)-(my*my).astype(f)).astype(f)
    sxy=((seqsum((x*y).astype(f))/f(9)).astype(f)-(mx*my).astype(f)).astype(f)
    n_=((f(2)*mx*my+C1).astype(f)*(f(2)*sxy+C2).astype(f)).astype(f)
    d_=(((mx*mx).astype(f)+(my*my).astype(f)+C1).astype(f)*(sx+sy+C2).astype(f)).astype(f)
    return np.clip(((f(1)-(n_/d_).astype(f))/f(2)).astype(f),0,1)
def v2(x,y):
    f=np.float32
    def sepsum(a):
        h=[(a[:,3*r]+a[:,3*r+1]).astype(f)+a[:,3*r+2] for r in range(3)]
        return ((h[1]+h[2]).astype(f)+h[0]).astype(f)
    sx=sepsum(x); sy=sepsum(y); sxx=sepsum((x*x).astype(f)); syy=sepsum((y*y).astype(f)); sxy=sepsum((x*y).astype(f))
    c1=f(81)*C1; c2=f(81)*C2
    p=(sx*sy).astype(f); A1=(f(2)*p+c1).astype(f); A2=(f(-2)*p+(f(18)*sxy+c2).astype(f)).astype(f)
    q=(sx*sx+sy*sy).astype(f); B1=(q+c1).astype(f); B2=((f(9)*(sxx+syy).astype(f)+c2).astype(f)-q).astype(f)
    s=((A1*A2).astype(f)/(B1*B2).astype(f)).astype(f)
    return np.clip((f(0.5)-f(0.5)*s).astype(f),0,1)
e=exact(x,y); r=ref32(x,y); v=v2(x,y)
for name,a in (("reference-order fp32",r),("grouped program formula",v)):
    err=np.abs(a.astype(np.float64)-e)
    print("%-26s abs err: median %.1e  99%% %.1e  max %.1e" % (name, np.median(err), np.quantile(err,0.99), err.max()))
print("ref vs grouped: median %.1e 99%% %.1e max %.1e" % tuple(f(np.abs(r.astype(np.float64)-v)) for f in (np.median, lambda a: np.quantile(a,0.99), np.max)))
