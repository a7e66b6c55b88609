"""RESEARCH PROTOTYPE (NumPy, CPU) -- not part of the product, not imported by anything.

A primal-dual method for the MVIE solves that dominate the C2 step (35 Newton iterations per solve with the
log-barrier path following of csrc/bp_mvie*.cuh).  Formulation that made Mehrotra's predictor-corrector robust:

  min -sum_k w_k log x_dk  s.t. c_i(x) = h_i + G_i x in Q^4      (w = 1, 2, 1 on the Cholesky pivots, quirk Q1)

is written WITHOUT a nonlinear objective: the pivots get dual variables y_k with the FIXED complementarity target
x_dk y_k = w_k (a weighted centre), the cones the usual c_i o z_i = mu e with mu -> 0, and stationarity is the linear
equation  sum_i G_i^T z_i + sum_k e_dk y_k = 0.  Only bilinear products remain, Nesterov-Todd scaling applies per
cone in closed form, and the reduced system is the same NV x NV matrix shape as the barrier Hessian (rank-one term
from the scaling point + the J-term with weight eta^-2), so the warp solver's column-dot machinery carries over.

Measured on the 256 C2 sets (this script, `python tools/research/mvie_primal_dual_prototype.py`):
  * cold start: 11-12 iterations when it converges, but jams at the boundary on ~10 % of the free-centre solves;
  * hybrid, robust on every set: centre with the existing barrier Newton at t = 1 (5.2 iterations), then primal-dual
    from that point: 9-10 iterations to gap 1e-9 (15 in total against 35);
  * the primal-dual x converges like sqrt(gap) (0.02 sqrt(gap) relative), the barrier's like gap: to keep today's
    1e-9 agreement with the oracle, finish with 2-3 barrier Newton steps at t = 2 m / gap (+ the last two stages).
The serial C++ specification (csrc/bp_mvie_pd.cuh, host-tested) measures 5 + 9 + 12 = 26 Newton iterations against 35:
the finishing barrier stages have to re-centre the primal-dual point."""
import numpy as np, sys, ctypes
sys.path.insert(0,'/root/repo')
from boundplanner_b200 import scenes
def jdet(v): n=np.sqrt(np.sum(v[:,1:]**2,axis=1)); return (v[:,0]-n)*(v[:,0]+n)
def Jm(v): o=v.copy(); o[:,1:]*=-1; return o
def jdot(a,b): return np.sum(a*b,axis=1)
def cone_of(A,b,c0,x,NV):
    cen=x[6:9] if NV==9 else c0
    c=np.zeros((A.shape[0],4))
    c[:,0]=b-A@cen
    c[:,1]=x[0]*A[:,0]+x[1]*A[:,1]+x[3]*A[:,2]
    c[:,2]=x[2]*A[:,1]+x[4]*A[:,2]
    c[:,3]=x[5]*A[:,2]
    return c
def Gt(A,p,NV):
    """sum_i G_i^T p_i for per-row 4-vectors p."""
    out=np.zeros(NV); a0,a1,a2=A[:,0],A[:,1],A[:,2]
    out[0]=p[:,1]@a0; out[1]=p[:,1]@a1; out[2]=p[:,2]@a1; out[3]=p[:,1]@a2; out[4]=p[:,2]@a2; out[5]=p[:,3]@a2
    if NV==9: out[6]=-(p[:,0]@a0); out[7]=-(p[:,0]@a1); out[8]=-(p[:,0]@a2)
    return out
def Gx(A,dx,NV):
    dc=np.zeros((A.shape[0],4))
    if NV==9: dc[:,0]=-(A@dx[6:9])
    dc[:,1]=dx[0]*A[:,0]+dx[1]*A[:,1]+dx[3]*A[:,2]
    dc[:,2]=dx[2]*A[:,1]+dx[4]*A[:,2]
    dc[:,3]=dx[5]*A[:,2]
    return dc
def max_step(c,dc):
    """largest alpha with c + alpha dc in the cone, per row -> min"""
    a=jdet_raw(dc); b=2*(c[:,0]*dc[:,0]-np.sum(c[:,1:]*dc[:,1:],axis=1)); cc=jdet(c)
    am=np.full(c.shape[0],np.inf)
    neg=dc[:,0]<0
    am[neg]=-c[neg,0]/dc[neg,0]
    disc=b*b-4*a*cc
    ok=disc>=0
    sq=np.sqrt(np.where(ok,disc,0)); q=-0.5*(b+np.where(b>=0,sq,-sq))
    with np.errstate(divide='ignore',invalid='ignore'):
        r1=np.where(a!=0,q/a,np.inf); r2=np.where(q!=0,cc/q,np.inf)
    lin=(np.abs(a)<1e-300)
    r1=np.where(lin,np.where(b<0,-cc/np.where(b<0,b,-1),np.inf),r1); r2=np.where(lin,np.inf,r2)
    for r in (r1,r2):
        sel=ok&(r>0)
        am[sel]=np.minimum(am[sel],r[sel])
    return am.min()
def jdet_raw(v): return v[:,0]**2-np.sum(v[:,1:]**2,axis=1)
def pd_solve(A,b,c0,NV,tol=1e-9,maxit=40,verbose=False,frac=0.99):
    m=A.shape[0]; D=[0,2,5]; w=np.array([1.0,2.0,1.0])
    nrm=np.linalg.norm(A,axis=1)
    r=0.5*np.min((b-A@c0)/nrm)
    x=np.zeros(NV); x[0]=x[2]=x[5]=r
    if NV==9: x[6:9]=c0
    c=cone_of(A,b,c0,x,NV)
    y=w/x[D]; z=Jm(c)/jdet(c)[:,None]
    e=np.zeros((m,4)); e[:,0]=1
    for it in range(1,maxit+1):
        c=cone_of(A,b,c0,x,NV)
        rd=Gt(A,z,NV); rd[D]+=y
        mu=float(np.sum(c*z))/m
        dc_=jdet(c); dz_=jdet(z)
        ct=c/np.sqrt(dc_)[:,None]; zt=z/np.sqrt(dz_)[:,None]
        gam=np.sqrt((1+jdot(ct,zt))/2)
        wb=(ct+Jm(zt))/(2*gam)[:,None]                 # scaling point, det = 1
        eta=(dc_/dz_)**0.25
        v=wb.copy(); v[:,0]+=1; v/=np.sqrt(2*v[:,0])[:,None]
        Jv=Jm(v)
        Wa=lambda u: eta[:,None]*(2*v*jdot(v,u)[:,None]-Jm(u))
        Wia=lambda u: (2*Jv*jdot(Jv,u)[:,None]-Jm(u))/eta[:,None]
        lam=Wa(z)
        # Hessian: sum eta^-2 (2 (G^T J wb)(G^T J wb)^T - G^T J G) + diag(y/x)
        q=Jm(wb)
        a0,a1,a2=A[:,0],A[:,1],A[:,2]
        rr=np.zeros((m,NV))
        rr[:,0]=q[:,1]*a0; rr[:,1]=q[:,1]*a1; rr[:,2]=q[:,2]*a1; rr[:,3]=q[:,1]*a2; rr[:,4]=q[:,2]*a2; rr[:,5]=q[:,3]*a2
        if NV==9: rr[:,6]=-q[:,0]*a0; rr[:,7]=-q[:,0]*a1; rr[:,8]=-q[:,0]*a2
        om=1/eta**2
        H=2*(rr*om[:,None]).T@rr
        Wm=(A*om[:,None]).T@A                           # sum om a a^T
        for g,cs in (([0,1,3],[0,1,2]),([2,4],[1,2]),([5],[2])):
            for i_,ci in zip(g,cs):
                for j_,cj in zip(g,cs): H[i_,j_]+=Wm[ci,cj]
        if NV==9: H[6:9,6:9]-=Wm
        H[D,D]+=y/x[D]
        def jdiv(l,d):
            y0=(l[:,0]*d[:,0]-np.sum(l[:,1:]*d[:,1:],axis=1))/jdet(l)
            yb=(d[:,1:]-y0[:,None]*l[:,1:])/l[:,0:1]
            return np.concatenate((y0[:,None],yb),axis=1)
        def jprod(a_,b_): return np.concatenate((jdot(a_,b_)[:,None], a_[:,0:1]*b_[:,1:]+b_[:,0:1]*a_[:,1:]),axis=1)
        def solve(ds_rhs, comp_rhs):
            t=Wia(jdiv(lam,ds_rhs))
            rhs=rd+Gt(A,t,NV); rhs[D]+=comp_rhs/x[D]
            dx=np.linalg.solve(H,rhs)
            dc=Gx(A,dx,NV)
            dz=t-Wia(Wia(dc))
            dy=comp_rhs/x[D]-(y/x[D])*dx[D]
            return dx,dc,dz,dy
        def step_len(dx,dc,dz,dy):
            am=min(max_step(c,dc),max_step(z,dz))
            for k,dk in enumerate(D):
                if dx[dk]<0: am=min(am,-x[dk]/dx[dk])
                if dy[k]<0: am=min(am,-y[k]/dy[k])
            return am
        ll=jprod(lam,lam)
        dxa,dca,dza,dya=solve(-ll, w-x[D]*y)
        aa=min(1.0,step_len(dxa,dca,dza,dya))
        sigma=(1-aa)**3
        dsc=-ll-jprod(Wia(dca),Wa(dza))+sigma*mu*e
        dx,dc,dz,dy=solve(dsc, w-x[D]*y-dxa[D]*dya)
        al=min(1.0,frac*step_len(dx,dc,dz,dy))
        x=x+al*dx; z=z+al*dz; y=y+al*dy
        c=cone_of(A,b,c0,x,NV)
        gap=float(np.sum(c*z))
        rdn=Gt(A,z,NV); rdn[D]+=y; rdn=np.linalg.norm(rdn)/max(1.0,np.linalg.norm(y))
        cw=np.abs(x[D]*y-w).max()
        if verbose: print(it,f"aa {aa:.3g} sigma {sigma:.2e} al {al:.3g} gap {gap:.3e} rd {rdn:.2e} cw {cw:.2e}")
        if not np.isfinite(gap): return x,99
        if gap<tol and rdn<1e-9 and cw<1e-9: break
    return x,it
def pd_from(A,b,c0,NV,x,mu0,tol=1e-9,verbose=False,maxit=40,frac=0.99):
    m=A.shape[0]; D=[0,2,5]; w=np.array([1.0,2.0,1.0])
    x=x.copy()
    c=cone_of(A,b,c0,x,NV)
    y=w/x[D]; z=2*mu0*Jm(c)/jdet(c)[:,None]
    e=np.zeros((m,4)); e[:,0]=1
    for it in range(1,maxit+1):
        c=cone_of(A,b,c0,x,NV)
        rd=Gt(A,z,NV); rd[D]+=y
        mu=float(np.sum(c*z))/m
        dc_=jdet(c); dz_=jdet(z)
        ct=c/np.sqrt(dc_)[:,None]; zt=z/np.sqrt(dz_)[:,None]
        gam=np.sqrt((1+jdot(ct,zt))/2)
        wb=(ct+Jm(zt))/(2*gam)[:,None]                 # scaling point, det = 1
        eta=(dc_/dz_)**0.25
        v=wb.copy(); v[:,0]+=1; v/=np.sqrt(2*v[:,0])[:,None]
        Jv=Jm(v)
        Wa=lambda u: eta[:,None]*(2*v*jdot(v,u)[:,None]-Jm(u))
        Wia=lambda u: (2*Jv*jdot(Jv,u)[:,None]-Jm(u))/eta[:,None]
        lam=Wa(z)
        # Hessian: sum eta^-2 (2 (G^T J wb)(G^T J wb)^T - G^T J G) + diag(y/x)
        q=Jm(wb)
        a0,a1,a2=A[:,0],A[:,1],A[:,2]
        rr=np.zeros((m,NV))
        rr[:,0]=q[:,1]*a0; rr[:,1]=q[:,1]*a1; rr[:,2]=q[:,2]*a1; rr[:,3]=q[:,1]*a2; rr[:,4]=q[:,2]*a2; rr[:,5]=q[:,3]*a2
        if NV==9: rr[:,6]=-q[:,0]*a0; rr[:,7]=-q[:,0]*a1; rr[:,8]=-q[:,0]*a2
        om=1/eta**2
        H=2*(rr*om[:,None]).T@rr
        Wm=(A*om[:,None]).T@A                           # sum om a a^T
        for g,cs in (([0,1,3],[0,1,2]),([2,4],[1,2]),([5],[2])):
            for i_,ci in zip(g,cs):
                for j_,cj in zip(g,cs): H[i_,j_]+=Wm[ci,cj]
        if NV==9: H[6:9,6:9]-=Wm
        H[D,D]+=y/x[D]
        def jdiv(l,d):
            y0=(l[:,0]*d[:,0]-np.sum(l[:,1:]*d[:,1:],axis=1))/jdet(l)
            yb=(d[:,1:]-y0[:,None]*l[:,1:])/l[:,0:1]
            return np.concatenate((y0[:,None],yb),axis=1)
        def jprod(a_,b_): return np.concatenate((jdot(a_,b_)[:,None], a_[:,0:1]*b_[:,1:]+b_[:,0:1]*a_[:,1:]),axis=1)
        def solve(ds_rhs, comp_rhs):
            t=Wia(jdiv(lam,ds_rhs))
            rhs=rd+Gt(A,t,NV); rhs[D]+=comp_rhs/x[D]
            dx=np.linalg.solve(H,rhs)
            dc=Gx(A,dx,NV)
            dz=t-Wia(Wia(dc))
            dy=comp_rhs/x[D]-(y/x[D])*dx[D]
            return dx,dc,dz,dy
        def step_len(dx,dc,dz,dy):
            am=min(max_step(c,dc),max_step(z,dz))
            for k,dk in enumerate(D):
                if dx[dk]<0: am=min(am,-x[dk]/dx[dk])
                if dy[k]<0: am=min(am,-y[k]/dy[k])
            return am
        ll=jprod(lam,lam)
        dxa,dca,dza,dya=solve(-ll, w-x[D]*y)
        aa=min(1.0,step_len(dxa,dca,dza,dya))
        sigma=(1-aa)**3
        dsc=-ll-jprod(Wia(dca),Wa(dza))+sigma*mu*e
        dx,dc,dz,dy=solve(dsc, w-x[D]*y-dxa[D]*dya)
        al=min(1.0,frac*step_len(dx,dc,dz,dy))
        x=x+al*dx; z=z+al*dz; y=y+al*dy
        c=cone_of(A,b,c0,x,NV)
        gap=float(np.sum(c*z))
        rdn=Gt(A,z,NV); rdn[D]+=y; rdn=np.linalg.norm(rdn)/max(1.0,np.linalg.norm(y))
        cw=np.abs(x[D]*y-w).max()
        if verbose: print(it,f"aa {aa:.3g} sigma {sigma:.2e} al {al:.3g} gap {gap:.3e} rd {rdn:.2e} cw {cw:.2e}")
        if not np.isfinite(gap): return x,99
        if gap<tol and rdn<1e-9 and cw<1e-9: break
    return x,it

if __name__=="__main__":
    d=np.load('/root/repo/gpurun_out/c2_sets.npz')
    boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2()
    hh=ctypes.CDLL('/root/repo/tests/_build/libbp_host_harness.so')
    P=ctypes.POINTER(ctypes.c_double); dp=lambda a:np.ascontiguousarray(a).ctypes.data_as(P)
    tol=float(sys.argv[1]) if len(sys.argv)>1 else 1e-9
    np.seterr(all='ignore')
    for NV in (6,9):
        its=[];errs=[]
        for s_ in range(0,256):
            m=int(d['m'][s_]); A=np.ascontiguousarray(d['A'][s_,:m]); b=np.ascontiguousarray(d['b'][s_,:m]); c=np.ascontiguousarray(seeds[s_])
            x,it=pd_solve(A,b,c,NV,tol=tol)
            E,L,cen,itc=np.zeros((3,3)),np.zeros(6),np.zeros(3),ctypes.c_int()
            hh.hh_mvie_ws(dp(A),dp(b),m,int(NV==9),dp(c),None,ctypes.c_double(1.0),dp(E),dp(L),dp(cen),ctypes.byref(itc))
            its.append(it); errs.append(np.abs(x[:6]-L).max()/np.abs(L).max())
        its=np.array(its); errs=np.array(errs)
        print(f"tol {tol:g} NV={NV}: iters mean {its[its<40].mean():.1f} max {its[its<40].max()} | not converged {int((its>=40).sum())} | err vs barrier: median {np.median(errs):.1e} max {errs.max():.1e}")
def dbg(NV, which=None):
    np.seterr(all='ignore')
    d=np.load('/root/repo/gpurun_out/c2_sets.npz')
    boxes, inflate, seeds, ws_min, ws_max = scenes.config_c2()
    bad=[]
    for s_ in range(0,256):
        m=int(d['m'][s_]); A=np.ascontiguousarray(d['A'][s_,:m]); b=np.ascontiguousarray(d['b'][s_,:m]); c=np.ascontiguousarray(seeds[s_])
        x,it=pd_solve(A,b,c,NV,tol=1e-8)
        if it>=40: bad.append(s_)
    print("bad", bad)
    s_=bad[0] if which is None else which
    m=int(d['m'][s_]); A=np.ascontiguousarray(d['A'][s_,:m]); b=np.ascontiguousarray(d['b'][s_,:m]); c=np.ascontiguousarray(seeds[s_])
    pd_solve(A,b,c,NV,tol=1e-8,verbose=True,maxit=25)

def barrier_center(A,b,c0,NV,t,x,tol=1e-2,maxit=40):
    """plain primal Newton on F_t = t f + sum -log det(c_i) from x (strictly feasible) until lambda^2 < tol"""
    D=[0,2,5]; w=np.array([1.0,2.0,1.0]); m=A.shape[0]
    its=0
    def F(xv):
        if np.any(xv[D]<=0): return np.inf
        c=cone_of(A,b,c0,xv,NV); dt=jdet(c)
        if np.any(dt<=0) or np.any(c[:,0]<=0): return np.inf
        return -t*np.sum(w*np.log(xv[D]))-np.sum(np.log(dt))
    for k in range(maxit):
        its+=1
        c=cone_of(A,b,c0,x,NV); psi=jdet(c)
        q=Jm(c)   # grad of det/2 direction
        a0,a1,a2=A[:,0],A[:,1],A[:,2]
        rr=np.zeros((m,NV))
        rr[:,0]=q[:,1]*a0; rr[:,1]=q[:,1]*a1; rr[:,2]=q[:,2]*a1; rr[:,3]=q[:,1]*a2; rr[:,4]=q[:,2]*a2; rr[:,5]=q[:,3]*a2
        if NV==9: rr[:,6]=-q[:,0]*a0; rr[:,7]=-q[:,0]*a1; rr[:,8]=-q[:,0]*a2
        # grad of -log psi = -(2/psi) * G^T J c ; note G^T(Jc) = rr with q=Jc
        g=-(2/psi)@rr
        H=(rr*(4/psi**2)[:,None]).T@rr
        Wm=(A*(2/psi)[:,None]).T@A
        for gi,cs in (([0,1,3],[0,1,2]),([2,4],[1,2]),([5],[2])):
            for i_,ci in zip(gi,cs):
                for j_,cj in zip(gi,cs): H[i_,j_]+=Wm[ci,cj]
        if NV==9: H[6:9,6:9]-=Wm
        g[D]-=t*w/x[D]; H[D,D]+=t*w/x[D]**2
        dx=np.linalg.solve(H,-g); lam2=-g@dx
        if lam2<tol: x=x+dx if F(x+dx)<np.inf else x; break
        al=1.0; F0=F(x)
        while not (F(x+al*dx)<=F0-0.25*al*lam2): al*=0.5
        x=x+al*dx
    return x,its
def hybrid(A,b,c0,NV,t0=20.0,tol=1e-9,verbose=False):
    m=A.shape[0]; D=[0,2,5]; w=np.array([1.0,2.0,1.0])
    nrm=np.linalg.norm(A,axis=1); r=0.5*np.min((b-A@c0)/nrm)
    x=np.zeros(NV); x[0]=x[2]=x[5]=r
    if NV==9: x[6:9]=c0
    nb=0
    x,k=barrier_center(A,b,c0,NV,1.0,x); nb+=k
    if t0>1: x,k=barrier_center(A,b,c0,NV,t0,x); nb+=k
    xpd,npd=pd_from(A,b,c0,NV,x,1.0/t0,tol,verbose)
    return xpd,nb,npd
