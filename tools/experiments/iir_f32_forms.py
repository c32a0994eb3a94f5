"""fp32 accuracy of three ways to evaluate one cascaded-biquad section, emulated in numpy (float32 values, fused multiply-add = one rounding of the\nexact product-sum) against the fp64 oracle: direct form (round 1), difference form with the running difference carried (D2, what the kernels do\nnow) and with the difference re-derived from v[n-1]-v[n-2] (D1 column = variant D1: c0*v[n-1], a2*d).  Writes the table profiles/r02_iir_f32_delta_form.txt was made from."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import oracle as O
f32 = np.float32
def r32(x): return np.asarray(x, dtype=np.float64).astype(np.float32).astype(np.float64)
def fma(a,b,c): return r32(a*b + c)   # a,b,c are fp32-valued float64 arrays
def add(a,b): return r32(a+b)
def mul(a,b): return r32(a*b)

def design(ftype, f0, fs, q=1.0, m=4):
    f = O.Iir(m, "generic", "port"); f.design(ftype, f0, fs, q)
    return f.coefficients()

def run_ref(coefs, x):
    out = np.empty_like(x)
    for c,(ft,f0,fs,q) in enumerate(coefs):
        f = O.Iir(4, "generic", "port"); f.design(ft,f0,fs,q)
        out[c] = f.process(x[c])
    return out

def run_direct(G, B, A, x):
    # G [C], B,A [C,m,3]; x [C,n] fp32-valued
    C, n = x.shape; m = B.shape[1]
    g = r32(G); b1 = r32(B[:,:,1]); b2 = r32(B[:,:,2]); na1 = r32(-A[:,:,1]); na2 = r32(-A[:,:,2])
    h = np.zeros((C, m+1, 2))
    y = np.empty_like(x)
    for i in range(n):
        in0 = mul(x[:,i], g); in1 = h[:,0,0].copy(); in2 = h[:,0,1].copy()
        h[:,0,1] = in1; h[:,0,0] = in0
        for j in range(m):
            v1 = h[:,j+1,0].copy(); v2 = h[:,j+1,1].copy()
            acc = fma(b2[:,j], in2, fma(b1[:,j], in1, in0))
            v = fma(na1[:,j], v1, fma(na2[:,j], v2, acc))
            h[:,j+1,1] = v1; h[:,j+1,0] = v
            in0, in1, in2 = v, v1, v2
        y[:,i] = in0
    return y

def run_delta(G, B, A, x, variant="D2"):
    C, n = x.shape; m = B.shape[1]
    g = r32(G); b1 = r32(B[:,:,1]); b2 = r32(B[:,:,2])
    c0 = r32(1.0 + A[:,:,1] + A[:,:,2])       # computed in fp64, rounded once
    gg = r32(-(1.0 + A[:,:,1]))               # a2 - c0
    a2 = r32(A[:,:,2])
    # state: v1, v2, d1 per section; input row history
    hin = np.zeros((C,2)); v1s = np.zeros((C,m)); v2s = np.zeros((C,m)); d1s = np.zeros((C,m))
    y = np.empty_like(x)
    for i in range(n):
        in0 = mul(x[:,i], g); in1 = hin[:,0].copy(); in2 = hin[:,1].copy()
        hin[:,1] = in1; hin[:,0] = in0
        for j in range(m):
            v1 = v1s[:,j].copy(); v2 = v2s[:,j].copy(); d1 = d1s[:,j].copy()
            acc = fma(b2[:,j], in2, fma(b1[:,j], in1, in0))
            if variant == "D2":
                t = fma(-c0[:,j], v2, acc)
                d = fma(gg[:,j], d1, t)
            else:
                t = fma(-c0[:,j], v1, acc)
                d = fma(a2[:,j], d1, t)
            v = add(v1, d)
            v2s[:,j] = v1; v1s[:,j] = v; d1s[:,j] = d
            in0, in1, in2 = v, v1, v2
        y[:,i] = in0
    return y

cfgs = [(1,200.,39000.,1.),(1,500.,100e3,1.),(1,1000.,100e3,1.),(1,10e3,100e3,1.),(1,2000.,39000.,1.),(1,15000.,39000.,1.),
        (2,200.,39000.,1.),(2,500.,100e3,1.),(2,1000.,100e3,1.),(2,10e3,100e3,1.),(2,15000.,39000.,1.),(2, 45e3, 100e3, 1.),
        (3,200.,39000.,1.4),(3,2000.,39000.,0.8),(3,15000.,39000.,2.0),(3,500.,100e3,1.1), (1, 100., 100e3, 1.), (1, 49e3, 100e3, 1.)]
G=[];B=[];A=[]
for c in cfgs:
    g,b,a = design(*c); G.append(g);B.append(b);A.append(a)
G=np.array(G);B=np.array(B);A=np.array(A)
n = int(sys.argv[1]) if len(sys.argv)>1 else 20000
rng = np.random.default_rng(1)
for name, x in (("impulse", np.tile(np.eye(1,n), (len(cfgs),1))), ("noise", r32(rng.standard_normal((len(cfgs), n))))):
    ref = run_ref(cfgs, x)
    pk = np.abs(ref).max(axis=1)
    res = {}
    for nm, fn in (("direct", lambda: run_direct(G,B,A,x)), ("D2", lambda: run_delta(G,B,A,x,"D2")), ("D1", lambda: run_delta(G,B,A,x,"D1"))):
        y = fn(); res[nm] = np.abs(y-ref).max(axis=1)/pk
    print(name)
    for i,c in enumerate(cfgs):
        print(f"  {c}: direct {res['direct'][i]:.2e}  D2 {res['D2'][i]:.2e}  D1 {res['D1'][i]:.2e}")
