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

# asymptotic figures: the same reductions / copy on large tensors (launch ramp and tail amortised)
for mb in (66, 132, 264, 1056, 4224):
    n = mb * 2 ** 20 // 4
    ys = [torch.rand(n, device=dev) for _ in range(2)]
    def t(f, reps=20):
        for i in range(3): f(ys[i % 2])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps): f(ys[i % 2])
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    us_sum, us_clone = t(lambda x: x.sum()), t(lambda x: x.clone())
    print('%5d MB  sum %8.1f us = %5.0f GB/s read   clone %8.1f us = %5.0f GB/s read+write' % (mb, us_sum, n * 4 / us_sum / 1e3, us_clone, 2 * n * 4 / us_clone / 1e3))
    del ys
