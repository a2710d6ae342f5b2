import sys, math, torch
sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/tests")
import kernels as K
DEV="cuda:0"
def randn(*s, seed=0, dtype=torch.float32, scale=1.0):
    g=torch.Generator(device=DEV).manual_seed(seed); return (torch.randn(*s, generator=g, device=DEV)*scale).to(dtype)
B,img,txt,N,Kd=1,8192,256,3072,512
s=K.seq(B,img,txt); rows=K.rows(s)
a=randn(rows,Kd,seed=80,dtype=torch.bfloat16); w=[randn(N,Kd,seed=81+i,dtype=torch.bfloat16,scale=1/math.sqrt(Kd)) for i in range(2)]; b=[randn(N,seed=83+i,scale=0.1) for i in range(2)]
gate=randn(B,2,6*N,seed=81); res0=randn(rows,N,seed=82)
sep=res0.clone()
K.gemm(s,a,w,b,sep,K.L.EPI_GATE_RESID_F32,gate=gate[:,:,2*N:],gate_bstride=12*N,gate_sstride=6*N,cta_group=2)
want=K.ln_modulate(s,sep,gate,12*N,6*N,3*N,4*N,N)
for it in range(3):
    fused=res0.clone(); xm=torch.full((rows,N),5.0,dtype=torch.bfloat16,device=DEV)
    ln=dict(out=xm,mod=gate,bstride=12*N,sstride=6*N,shift_off=3*N,scale_off=4*N)
    K.gemm(s,a,w,b,fused,K.L.EPI_GATE_RESID_F32,gate=gate[:,:,2*N:],gate_bstride=12*N,gate_sstride=6*N,cta_group=2,ln=ln)
    torch.cuda.synchronize()
    d=(xm.float()-want.float()).abs()
    bad=(d>0)
    print("iter",it,"resid equal",torch.equal(fused,sep),"mismatch elems",int(bad.sum()),"rows",int(bad.any(1).sum()),"max diff",d.max().item(), "rows idx", bad.any(1).nonzero().flatten()[:10].tolist())
    # ulp-level?
    rel=(d/(want.float().abs()+1e-3)).max().item(); print("  max rel",rel)
