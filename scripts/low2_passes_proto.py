"""Numpy check of the bit-reversed DIT/DIF pass formulas used by tree_low2.cuh."""
import numpy as np
def bitrev(x,bits):
    r=0
    for i in range(bits):
        r=(r<<1)|((x>>i)&1)
    return r
def inv_pass(A,N,R,s):
    r=R.bit_length()-1
    for g in range(N//(R*s)):
        for o in range(s):
            base=g*R*s+o
            v=np.array([A[base+bitrev(q,r)*s]*np.exp(2j*np.pi*q*o/(R*s)) for q in range(R)])
            y=np.array([sum(v[q]*np.exp(2j*np.pi*q*n/R) for q in range(R)) for n in range(R)])
            for n in range(R): A[base+n*s]=y[n]
def fwd_pass(A,N,R,s):
    r=R.bit_length()-1
    for g in range(N//(R*s)):
        for o in range(s):
            base=g*R*s+o
            v=np.array([A[base+n*s] for n in range(R)])
            y=np.array([sum(v[n]*np.exp(-2j*np.pi*q*n/R) for n in range(R))*np.exp(-2j*np.pi*q*o/(R*s)) for q in range(R)])
            for q in range(R): A[base+bitrev(q,r)*s]=y[q]
rng=np.random.default_rng(0)
for N,plan in ((16,[4,4]),(32,[4,8]),(64,[4,16]),(128,[4,4,8]),(512,[4,8,16])):
    x=rng.normal(size=N)+1j*rng.normal(size=N)
    X=np.fft.fft(x)
    bits=N.bit_length()-1
    A=np.array([X[bitrev(p,bits)] for p in range(N)])
    s=1
    for R in plan:
        inv_pass(A,N,R,s); s*=R
    print(N,'inv err',abs(A/N-x).max())
    # forward: reverse plan order, strides descending
    A=x.copy()
    s=N
    for R in reversed(plan):
        s//=R
        fwd_pass(A,N,R,s)
    print(N,'fwd err',abs(A-np.array([X[bitrev(p,bits)] for p in range(N)])).max())
