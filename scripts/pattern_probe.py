"""How fast can the class tensor be read at all in the operator's access pattern?  torch reductions over (B, C, A)."""
import torch
dev = torch.device('cuda', 0)
B, C, A = 32, 21, 24564
xs = [torch.rand(B, C, A, device=dev) for _ in range(4)]
def timeit(f, n=200):
    for i in range(20): f(xs[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): f(xs[i % 4])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
nbytes = B * C * A * 4
for name, f in (('amax over classes (B,C,A)->(B,A)', lambda x: x.amax(dim=1)),
                ('sum over everything', lambda x: x.sum()),
                ('copy', lambda x: x.clone())):
    us = timeit(f)
    print('%-36s %6.1f us  %5.0f GB/s read' % (name, us, nbytes / us / 1e3))
