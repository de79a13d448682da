import torch, time
torch.cuda.set_device(0)
n = 1228800000  # 4.9 GB of f32
a = torch.empty(n, dtype=torch.float32, device='cuda')
b = torch.empty(n // 4, dtype=torch.float32, device='cuda')
c = torch.empty(n // 4, dtype=torch.float32, device='cuda')
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: a.fill_(1.5)); print('fill  f32 4.9GB: %.3f ms  %.0f GB/s' % (ms, n * 4 / ms / 1e6))
ms = t(lambda: a.zero_()); print('zero  (memset)  : %.3f ms  %.0f GB/s' % (ms, n * 4 / ms / 1e6))
ms = t(lambda: c.copy_(b)); print('copy  1.2GB->1.2GB: %.3f ms  %.0f GB/s (r+w)' % (ms, 2 * (n // 4) * 4 / ms / 1e6))
# mixed: 24% read / 76% write emulation: read b (1.2GB), write a[:3*len(b)] 
v = a[: 3 * (n // 4)].view(3, -1)
ms = t(lambda: torch.add(b.unsqueeze(0), 1.0, out=None) if False else v.copy_(b.unsqueeze(0).expand(3, -1)))
print('read 1x, write 3x : %.3f ms  %.0f GB/s (r+w)' % (ms, 4 * (n // 4) * 4 / ms / 1e6))
