"""Minimax fit of the logistic-polynomial GELU used by the GEMM epilogues (csrc/epilogue_math.cuh)."""
import numpy as np
from scipy.special import erf
from scipy.optimize import least_squares
x = np.linspace(-12, 12, 60001)
Phi = 0.5*(1+erf(x/np.sqrt(2)))
gelu = x*Phi
UMAX=36.0
def model(c, x, dt=np.float64):
    x=x.astype(dt); u = np.minimum(x*x, dt(UMAX))
    p = x*(dt(c[0]) + u*(dt(c[1]) + u*dt(c[2])))
    return x/(1+np.exp(-p))
c0=[1.5950157548118395, 0.07401130356060703, -0.0007030353063620077]
w=np.ones_like(x)
for it in range(80):
    r = least_squares(lambda c: w*(model(c,x)-gelu), c0, xtol=1e-15, ftol=1e-15); c0=r.x
    e=np.abs(model(c0,x)-gelu); w=w*(1+2*e/e.max()); w/=w.mean()
print(list(c0), e.max(), x[e.argmax()])
e32=np.abs(model(c0,x,np.float32).astype(np.float64)-gelu); print("f32 err",e32.max())
# derivative of the model vs true derivative
u=np.minimum(x*x,UMAX); p=x*(c0[0]+u*(c0[1]+u*c0[2])); s=1/(1+np.exp(-p))
dp=np.where(x*x<UMAX, c0[0]+3*c0[1]*u+5*c0[2]*u*u, c0[0]+u*(c0[1]+u*c0[2]))
dm=s+x*s*(1-s)*dp
dt=Phi+x*np.exp(-x*x/2)/np.sqrt(2*np.pi)
print("deriv err", np.abs(dm-dt).max(), x[np.abs(dm-dt).argmax()])
L=np.log2(np.e)
print("K (=-c*log2e):", [-c*L for c in c0])
print("poly at umax", c0[0]+UMAX*(c0[1]+UMAX*c0[2]))
