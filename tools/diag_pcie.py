import torch, time
def t(fn, n=5):
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
d = torch.empty(2112861200//4, dtype=torch.float32, device='cuda')
h = torch.empty(d.shape, dtype=d.dtype, pin_memory=True)
h2 = torch.empty(536870912, dtype=torch.int8).pin_memory()
d2 = torch.empty_like(h2, device='cuda')
h.copy_(d); d2.copy_(h2)
print('d2h 2.1GB ms', t(lambda: h.copy_(d, non_blocking=True)))
print('h2d 0.54GB ms', t(lambda: d2.copy_(h2, non_blocking=True)))
s = torch.cuda.Stream()
def both():
    d2.copy_(h2, non_blocking=True)
    with torch.cuda.stream(s):
        h.copy_(d, non_blocking=True)
def fin():
    torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): both()
fin(); e1.record(); torch.cuda.synchronize(); print('overlapped ms', e0.elapsed_time(e1)/5)
