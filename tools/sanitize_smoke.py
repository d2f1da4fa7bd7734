import sys, importlib, numpy as np
sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/tests")
fb=importlib.import_module("faldoi-ipol_b200")
g=dict(np.load("/root/repo/tests/golden/crop_b.npz"))
for m,w,it in ((0,1,400),(4,1,400),(2,1,400),(6,1,400),(8,1,3)):
    p=fb.default_params(m,it,w); 
    if m!=8: p.max_iters=20
    u,chi,its,errs=fb.global_solve(m,g["I0n"],g["I1n"],g["u0"],Im1=g["Im1n"],lab=g["lab"],chi=g["chi0"] if m==8 else None,params=p)
    print(m,its,float(np.abs(u).max()))
s=fb.Stripes(61,45,[0,0,0]); s.upload(g["I0n"],g["I1n"],g["u0"]); p=fb.default_params(0,400,1); p.max_iters=10; s.run(p); print(s.download()[0].shape)
