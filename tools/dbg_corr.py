import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from master_thesis_b200 import ops, _lib
torch.manual_seed(0)
for (B,C,F,P,kind) in [(1,32,1,128,'ones'),(1,32,1,128,'rand'),(1,64,1,128,'rand'),(1,512,2,256,'rand')]:
    h=w=int(P**0.5) if int(P**0.5)**2==P else None
    if h is None: h, w = 8, P//8
    if kind=='ones':
        ft=torch.ones(B,C,h,w,device='cuda'); fr=torch.ones(B,C,F,h,w,device='cuda')
    else:
        ft=torch.rand(B,C,h,w,device='cuda'); fr=torch.rand(B,C,F,h,w,device='cuda')
    out=ops.corr4d(ft,None,fr,None)
    torch.cuda.synchronize()
    ws=[v for k,v in ops._workspaces.items() if k[0]=='corr'][0]
    sa=ws[:B*P*4].view(torch.float32)
    a=ft.reshape(B,C,P); bm=fr.reshape(B,C,F,P)
    an=a/(a.norm(dim=1,keepdim=True)+1e-9); bn=bm/(bm.norm(dim=1,keepdim=True)+1e-9)
    ref=torch.einsum('bkm,bkfn->bfmn',an,bn)
    o=out.reshape(B,F,P,P)
    print(kind,B,C,F,P,'tc=',_lib.load().mt_corr4d_uses_tensor_cores(C,P),'sa[:4]',sa[:4].tolist(),'exp',(1/(a.norm(dim=1)+1e-9))[0,:4].tolist())
    print('   out min/max',o.min().item(),o.max().item(),'ref min/max',ref.min().item(),ref.max().item(),'maxerr',(o-ref).abs().max().item())
    if (o-ref).abs().max().item()>1e-2:
        d=(o-ref).abs()[0,0]
        print('   nonzero frac', (o!=0).float().mean().item(), 'err rows', d.max(dim=1)[0][:8].tolist(), 'err cols', d.max(dim=0)[0][:8].tolist())
        print('   o[0,0,:4,:4]', o[0,0,:4,:4].tolist()); print('   r[0,0,:4,:4]', ref[0,0,:4,:4].tolist())
