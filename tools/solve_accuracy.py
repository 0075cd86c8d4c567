import numpy as np, mpmath
mpmath.mp.dps=50
rng=np.random.default_rng(0)
def kernel_inv(A):
    A=A.copy(); k=A.shape[0]; inv=np.ones(k)
    for p in range(k):
        col=np.where(np.arange(k)<p, -A[:,p]*inv, A[:,p]); d=col[p]; iv=1.0/d; inv[p]=iv
        f=A[:,p]*iv; f[p]=0; nA=A-np.outer(f,col); nA[:,p]=-f; nA[p,p]=1.0; A=nA
    return A*inv[:,None]
def sweep_inv(A):
    A=A.copy(); k=A.shape[0]
    for p in range(k):
        d=A[p,p]; col=A[:,p].copy()
        A=A-np.outer(col,col)/d; A[:,p]=col/d; A[p,:]=col/d; A[p,p]=-1/d
    return -A
def run(make, label, trials=40):
    errs={"full":[], "upper":[], "lower":[], "sweep":[], "lu":[], "chol":[]}
    for t in range(trials):
        A,y=make()
        k=A.shape[0]
        Am=mpmath.matrix(A.tolist()); ym=mpmath.matrix(y.tolist())
        qt=float((ym.T*(Am**-1)*ym)[0])
        Mi=kernel_inv(A)
        U=np.triu(Mi)+np.triu(Mi,1).T; Lw=np.tril(Mi)+np.tril(Mi,-1).T
        S=sweep_inv(A); LU=np.linalg.inv(A)
        Lc=np.linalg.cholesky(A); u=np.linalg.solve(Lc,y)
        for nm,q in (("full",y@Mi@y),("upper",y@U@y),("lower",y@Lw@y),("sweep",y@S@y),("lu",y@LU@y),("chol",u@u)):
            errs[nm].append(abs(q-qt)/abs(qt))
    print(label," ".join("%s %.1e"%(nm,np.max(v)) for nm,v in errs.items()))
k=16
def rank1():
    B=rng.standard_normal((40,k)); v=rng.standard_normal(k)
    A=B.T@B+0.25*np.eye(k)+1e6*np.outer(v,v); y=1e3*v*rng.standard_normal()+B.T@rng.standard_normal(40)
    return A,y
def rank3():
    B=rng.standard_normal((40,k)); V=rng.standard_normal((3,k))*np.array([1e3,1e2,1e1])[:,None]
    A=B.T@B+0.01*np.eye(k)+V.T@V; y=V.T@rng.standard_normal(3)+B.T@rng.standard_normal(40)
    return A,y
def graded():
    B=rng.standard_normal((40,k))*np.logspace(0,3,k)[None,:]
    A=B.T@B+0.25*np.eye(k); y=B.T@rng.standard_normal(40)
    return A,y
def benign():
    B=rng.standard_normal((160,k)); A=B.T@B+0.01*np.eye(k); y=B.T@rng.standard_normal(160)
    return A,y
def randy():
    B=rng.standard_normal((40,k)); v=rng.standard_normal(k)
    A=B.T@B+0.25*np.eye(k)+1e6*np.outer(v,v); y=rng.standard_normal(k)
    return A,y
for f in (benign,rank1,rank3,graded,randy): run(f,f.__name__)

print("---- sqrt-symmetric deferred sweep")
def sqrtsym_inv(A):
    A=A.copy(); k=A.shape[0]; inv=np.ones(k); piv=np.zeros(k,bool)
    for p in range(k):
        d=A[p,p]; rinv=1.0/np.sqrt(d)
        own=A[:,p]*rinv
        pub=np.where(piv, own*inv, own)
        nA=A-np.outer(own,pub)
        nA[:,p]=own*rinv
        nA[p,:]=A[p,:]          # deferred: pivot row untouched
        nA[p,p]=-1.0
        A=nA; inv[p]=rinv*rinv; piv[p]=True
    return -(A*inv[:,None])
def run2(make,label,trials=40):
    errs={"sq_full":[], "sq_upper":[], "sq_lower":[],"sq_sym":[], "sweep":[], "asym":[]}
    for t in range(trials):
        A,y=make(); 
        Am=mpmath.matrix(A.tolist()); ym=mpmath.matrix(y.tolist())
        qt=float((ym.T*(Am**-1)*ym)[0])
        Mi=sqrtsym_inv(A)
        U=np.triu(Mi)+np.triu(Mi,1).T; Lw=np.tril(Mi)+np.tril(Mi,-1).T
        S=sweep_inv(A)
        for nm,q in (("sq_full",y@Mi@y),("sq_upper",y@U@y),("sq_lower",y@Lw@y),("sq_sym",y@(0.5*(Mi+Mi.T))@y),("sweep",y@S@y)):
            errs[nm].append(abs(q-qt)/abs(qt))
        errs["asym"].append(np.max(np.abs(Mi-Mi.T))/np.max(np.abs(Mi)))
    print(label," ".join("%s %.1e"%(nm,np.max(v)) for nm,v in errs.items()))
for f in (benign,rank1,rank3,graded,randy): run2(f,f.__name__)

print("---- standard GJ (row broadcast, immediate scaling) and 8x8-block GJ")
def std_gj(A):
    A=A.copy(); k=A.shape[0]
    for p in range(k):
        d=A[p,p]; rowp=A[p,:]/d; colp=A[:,p].copy()
        A=A-np.outer(colp,rowp); A[:,p]=-colp/d; A[p,:]=rowp; A[p,p]=1/d
    return A
def blk_gj(A,b=8):
    A=A.copy(); k=A.shape[0]
    for s in range(0,k,b):
        P=np.linalg.inv(A[s:s+b,s:s+b])   # stands for the in-register 8x8 GJ
        R=A[s:s+b,:].copy(); Cc=A[:,s:s+b].copy()
        Lm=-Cc@P
        newA=A+Lm@R
        newA[:,s:s+b]=Lm
        newA[s:s+b,:]=P@R
        newA[s:s+b,s:s+b]=P
        A=newA
    return A
def run3(make,label,trials=40):
    errs={"std":[], "blk8":[], "sqrtsym":[]}
    for t in range(trials):
        A,y=make(); Am=mpmath.matrix(A.tolist()); ym=mpmath.matrix(y.tolist())
        qt=float((ym.T*(Am**-1)*ym)[0])
        for nm,Mi in (("std",std_gj(A)),("blk8",blk_gj(A)),("sqrtsym",sqrtsym_inv(A))):
            errs[nm].append(abs(y@Mi@y-qt)/abs(qt))
    print(label," ".join("%s %.1e"%(nm,np.max(v)) for nm,v in errs.items()))
for f in (benign,rank1,rank3,graded,randy): run3(f,f.__name__)

print("---- kernel variant with the REAL pivot row (deferred scaling)")
def kernel_realrow(A):
    A=A.copy(); k=A.shape[0]; inv=np.ones(k)
    for p in range(k):
        row=A[p,:].copy(); d=row[p]; iv=1.0/d; inv[p]=iv
        f=A[:,p]*iv; f[p]=0; nA=A-np.outer(f,row); nA[:,p]=-f; nA[p,p]=1.0; A=nA
    return A*inv[:,None]
def run4(make,label,trials=40):
    errs={"realrow":[], "sqrtsym":[]}
    for t in range(trials):
        A,y=make(); Am=mpmath.matrix(A.tolist()); ym=mpmath.matrix(y.tolist())
        qt=float((ym.T*(Am**-1)*ym)[0])
        for nm,Mi in (("realrow",kernel_realrow(A)),("sqrtsym",sqrtsym_inv(A))):
            errs[nm].append(abs(y@Mi@y-qt)/abs(qt))
    print(label," ".join("%s %.1e"%(nm,np.max(v)) for nm,v in errs.items()))
for f in (benign,rank1,rank3,graded,randy): run4(f,f.__name__)
