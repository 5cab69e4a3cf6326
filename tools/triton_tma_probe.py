import torch, triton, triton.language as tl
@triton.jit
def k(desc_in, out_ptr, BM: tl.constexpr, BN: tl.constexpr):
    t = desc_in.load([0, 0])
    offs = tl.arange(0, BM)[:, None] * BN + tl.arange(0, BN)[None, :]
    tl.store(out_ptr + offs, t)
from triton.tools.tensor_descriptor import TensorDescriptor
a = torch.arange(128*128, device="cuda", dtype=torch.float32).reshape(128,128)
out = torch.empty(16*32, device="cuda", dtype=torch.float32)
d = TensorDescriptor.from_tensor(a, [16, 32])
k[(1,)](d, out, 16, 32)
torch.cuda.synchronize()
print("triton tma ok", torch.equal(out.reshape(16,32), a[:16,:32]))
