"""One-off wider fuzz run on a GPU box: seeds 16..63 of both generators (tools/make_fuzz.py,
tools/make_fuzz_ico.py) compared bit for bit with the oracle.  The IIR files and modules must have been
generated and compiled in the build container first (see the header of DESIGN.md section 6); failing
seeds become committed fixtures.  Test infrastructure: uses oracle/ as the checker."""
import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tests/)
sys.path.insert(0,ROOT); sys.path.insert(0,ROOT+'/tests')
import numpy as np, torch, ctypes, traceback
import dawn_b200
from oracle import naive_interp
from util import HALO, bit_equal, iir_path, run_oracle
from dawn_b200.trimesh import TriMesh
LOC={"cells":0,"edges":1,"vertices":2}
fails=[]
def cart(name, opts):
    mod=dawn_b200.load_stencil(name, **opts)
    for (I,J,K) in ((17,11,7),(70,45,9)):
        ni,nj=I+2*HALO,J+2*HALO
        rng=np.random.default_rng(100)
        f={}
        for fd in mod.fields:
            d=fd["dims"]; f[fd["name"]]=rng.uniform(0.5,1.5,(K+1 if "k" in d else 1, nj if "j" in d else 1, ni if "i" in d else 1))
        g={"alpha":0.375} if "alpha" in mod.info.get("globals",[]) else None
        want=run_oracle(name,(ni,nj,K),f,globals_=g)
        dom=dawn_b200.Domain(ni,nj,K); dom.set_halos(HALO,HALO,HALO,HALO,0,0)
        st=mod(dom)
        for k,v in (g or {}).items(): st.set_global(k,v)
        stor=[dawn_b200.Storage(ni,nj,f[fd["name"]].shape[0],fd["dims"]).from_numpy(f[fd["name"]]) for fd in mod.fields]
        st.run(*stor); torch.cuda.synchronize()
        for fd,s in zip(mod.fields,stor):
            if not bit_equal(s.to_numpy(), want[fd["name"]]): fails.append((name,opts,fd["name"],(I,J,K)))
def ico(name):
    mod=dawn_b200.load_stencil(name)
    for (nx,ny,k) in ((9,7,5),(20,13,9)):
        m=TriMesh(nx,ny); rng=np.random.default_rng(200)
        f={}
        for fd in mod.fields:
            shape=[]
            if fd.get("k",1): shape.append(k)
            if fd["location"]!="none": shape.append(m.num(LOC[fd["location"]]))
            if fd["sparse"]: shape.append(fd["sparse"])
            f[fd["name"]]=rng.uniform(0.5,1.5,shape)
        g={"beta":0.625} if "beta" in mod.info.get("globals",[]) else None
        want=naive_interp.run_unstructured(iir_path(name),m,k,{n:v.copy() for n,v in f.items()},globals_=g)
        st=mod.on_mesh(m,k)
        for gn,gv in (g or {}).items():
            fn=getattr(mod.lib,"set_%s_%s"%(name,gn)); fn.argtypes=[ctypes.c_void_p,ctypes.c_double]; fn(st.handle,float(gv))
        tens=[]
        for fd in mod.fields:
            a=f[fd["name"]]
            if fd["sparse"]: a=np.ascontiguousarray(np.swapaxes(a,-1,-2))
            tens.append(torch.from_numpy(np.ascontiguousarray(a)).cuda())
        st.run(*tens); torch.cuda.synchronize()
        for fd,t in zip(mod.fields,tens):
            a=t.cpu().numpy(); a=np.swapaxes(a,-1,-2) if fd["sparse"] else a
            n=fd["name"]
            if fd["location"]=="none":
                ok=bit_equal(a,want[n])
            else:
                v=m.valid(LOC[fd["location"]]); ax=1 if fd.get("k",1) else 0
                ok=bit_equal(np.compress(v,a,axis=ax),np.compress(v,want[n],axis=ax))
            if not ok: fails.append((name,n,(nx,ny,k)))
for n in range(16,64):
    for opts in (dict(), dict(use_tma=0, inline_temporaries=0)):
        try: cart('fuzz_%02d'%n, opts)
        except Exception as e: fails.append(('fuzz_%02d'%n, opts, 'EXC', repr(e)[:120]))
    try: ico('fuzzico_%02d'%n)
    except Exception as e: fails.append(('fuzzico_%02d'%n, 'EXC', repr(e)[:120]))
print("FAILS", len(fails))
for f in fails[:30]: print(f)
