import sys; sys.path.insert(0,'/root/repo')
import torch
from x2i_b200.controlnext import ControlNeXtModel
from x2i_b200.flux import init_synthetic_
net = ControlNeXtModel().to('cuda', torch.bfloat16).eval()
init_synthetic_(net, seed=1, std=0.05)
hint = (torch.rand(1,3,1024,1024, device='cuda')*2-1).bfloat16()
t = torch.tensor([700.0], device='cuda')
x = torch.zeros(1,4096,3072, device='cuda', dtype=torch.bfloat16)
with torch.no_grad():
    net.forward_tokens(hint, t, add_to=x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    net.forward_tokens(hint, t, add_to=x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
