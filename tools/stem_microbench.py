import torch, sys
sys.path.insert(0,'.')
from spair_pytorch_b200 import ops
torch.backends.cudnn.allow_tf32=False
torch.backends.cudnn.benchmark=True
dev='cuda'
x=torch.rand(256,1,128,128,device=dev)
conv=torch.nn.Conv2d(1,128,4,3).to(dev)
pad=(9,14,9,14)
def lib():
    y=torch.relu(conv(torch.nn.functional.pad(x,pad)))
    return y
def fused():
    return ops.StemConvFunction.apply(x,conv.weight,conv.bias,3,9,9,50,50)
flush=torch.empty(64*1024*1024,device=dev)
def timeit(f,bwd):
    ts=[]
    for i in range(8):
        flush.fill_(1.0)
        if bwd:
            y=f(); g=torch.ones_like(y); torch.cuda.synchronize()
            flush.fill_(1.0)
            e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            e0.record(); y.backward(g); e1.record(); e1.synchronize()
        else:
            e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            e0.record(); y=f(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts)//2]
for name,f in (('library',lib),('fused',fused)):
    print(name,'fwd %.3f ms'%timeit(f,False),'bwd %.3f ms'%timeit(f,True))
out_bytes=256*128*50*50*4
print('fwd algorithmic bytes', out_bytes+x.numel()*4, 'bwd', 2*out_bytes+x.numel()*4)
