import sys, time, warnings; sys.path.insert(0,'/root/repo')
import torch, encodec_pytorch_b200 as E
from encodec_pytorch_b200.quantization.core_vq import kmeans
torch.manual_seed(0)
x = torch.randn(48000,128,device='cuda')
for it in (50,):
    torch.cuda.synchronize(); t=time.time()
    m,b = kmeans(x,1024,it); torch.cuda.synchronize(); print('kmeans',it,'iters', (time.time()-t)*1e3,'ms', int(b.sum()))
    torch.cuda.synchronize(); t=time.time()
    m,b = kmeans(x,1024,it); torch.cuda.synchronize(); print('kmeans',it,'iters', (time.time()-t)*1e3,'ms', int(b.sum()))
