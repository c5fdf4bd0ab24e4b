import torch, time, numpy as np
a = torch.empty(100 * 2**20, dtype=torch.uint8).pin_memory()
v = torch.from_numpy(a.numpy())
print('from_numpy(view of pinned).is_pinned()', v.is_pinned())
d = torch.empty_like(a, device='cuda')
for name, src, dst in (('H2D pinned', a, d), ('H2D via numpy view', v, d), ('D2H pinned', d, a)):
    for _ in range(2): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(name, '%.2f ms  %.1f GB/s' % (dt * 1e3, a.numel() / dt / 1e9))
p = torch.empty(100 * 2**20, dtype=torch.uint8)
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(p); torch.cuda.synchronize(); print('H2D pageable %.1f GB/s' % (p.numel() / (time.perf_counter() - t0) / 1e9))
import subprocess; print(subprocess.run(['nvidia-smi', '--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max', '--format=csv'], capture_output=True, text=True).stdout)
