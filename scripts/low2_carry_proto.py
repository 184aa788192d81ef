"""Numpy model of the spectrum-carry pair product of tree_low2.cuh (even bins = product values,\nodd bins = FFT_N(c_i w_2N^i) - c_N)."""
import numpy as np
rng=np.random.default_rng(1)
kappa=1
def sharp(f): return np.conj(f[::-1])
def prod(A,B):
    aA,bA=A; aB,bB=B
    a=np.convolve(aA,aB)-kappa*np.convolve(bA,sharp(bB))
    b=np.convolve(aA,bB)+np.convolve(bA,sharp(aB))
    return a,b
def evalN(p,N):  # index-space values at N-th roots, forward sign
    k=np.arange(N)[:,None]; i=np.arange(len(p))[None,:]
    return (p[None,:]*np.exp(-2j*np.pi*k*i/N)).sum(1)
d=8
mats=[(rng.normal(size=d+1)+1j*rng.normal(size=d+1), rng.normal(size=d+1)+1j*rng.normal(size=d+1)) for _ in range(4)]
# level state: per matrix V_a,V_b (N=2d points), top_a, top_b, bot_a, bot_b
def init(m,d):
    a,b=m; N=2*d
    return dict(Va=evalN(a,N),Vb=evalN(b,N),ta=a[d],tb=b[d],ba=a[0],bb=b[0],N=N)
def step(A,B):
    N=A['N']; k=np.arange(N); sg=(-1.0)**k
    Ca=A['Va']*B['Va']-kappa*sg*A['Vb']*np.conj(B['Vb'])
    Cb=A['Va']*B['Vb']+sg*A['Vb']*np.conj(B['Va'])
    cta=A['ta']*B['ta']-kappa*A['tb']*np.conj(B['bb'])
    ctb=A['ta']*B['tb']+A['tb']*np.conj(B['ba'])
    boa=A['ba']*B['ba']-kappa*A['bb']*np.conj(B['tb'])
    bob=A['ba']*B['bb']+A['bb']*np.conj(B['ta'])
    out={}
    coefs=[]
    for C,ct,bo,nm in ((Ca,cta,boa,'Va'),(Cb,ctb,bob,'Vb')):
        c=np.fft.fft(C)[(-np.arange(N))%N]/N   # inverse with + sign, /N  (ifft = conj-sign): use np.fft.ifft
        c=np.fft.ifft(C)    # numpy ifft: (1/N) sum C[k] e^{+2pi i k n/N}  matches inverse of forward e^{-}
        c=c.copy(); c[0]=bo
        y=c*np.exp(-2j*np.pi*np.arange(N)/(2*N))
        odd=np.fft.fft(y)-ct
        V=np.empty(2*N,complex); V[0::2]=C; V[1::2]=odd
        out[nm]=V
        coefs.append(np.concatenate([c,[ct]]))
    out.update(ta=cta,tb=ctb,ba=boa,bb=bob,N=2*N)
    return out,coefs
S=[init(m,d) for m in mats]
P01,c01=step(S[0],S[1]); P23,c23=step(S[2],S[3])
r01=prod(mats[0],mats[1]); r23=prod(mats[2],mats[3])
print('coef err',abs(c01[0]-r01[0]).max(),abs(c01[1]-r01[1]).max())
print('V err',abs(P01['Va']-evalN(r01[0],4*d)).max(),abs(P01['Vb']-evalN(r01[1],4*d)).max())
P,c=step(P01,P23)
r=prod(r01,r23)
print('lvl2 coef err',abs(c[0]-r[0]).max()/abs(r[0]).max(),abs(c[1]-r[1]).max()/abs(r[1]).max())
