"""Dev tool: fused vertex-feature front (ptk_vertex_front_fwd) vs the unfused path at the config-3 shape."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ptk_b200
dev=torch.device('cuda')
def timeit(fn, iters, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
torch.backends.cuda.matmul.allow_tf32 = False
enc, menc = ptk_b200.Positional_Encoder(448).to(dev), ptk_b200.Mask_Encoder(448).to(dev)
enc_u, menc_u = copy.deepcopy(enc), copy.deepcopy(menc); enc_u.fused=False
Bs=16
pos=(torch.rand(Bs,1949,3,device=dev)-0.5).requires_grad_(True); mask=torch.randint(0,4,(Bs,1949,1),device=dev).float()
img=torch.rand(Bs,1949,448,device=dev); w=torch.rand(Bs,1949,448,device=dev)
def front(e,m,train):
    def fn():
        if train:
            pos.grad=None; e.zero_grad(set_to_none=True); m.zero_grad(set_to_none=True)
            (ptk_b200.encoders.vertex_features(e,m,pos,mask,img)*w).sum().backward()
        else:
            with torch.no_grad(): ptk_b200.encoders.vertex_features(e,m,pos,mask,img)
    return fn
print("fused fwd %.3f ms, fwd+bwd %.3f | unfused fwd %.3f, fwd+bwd %.3f" % (timeit(front(enc,menc,False),20), timeit(front(enc,menc,True),10), timeit(front(enc_u,menc_u,False),20), timeit(front(enc_u,menc_u,True),10)))
