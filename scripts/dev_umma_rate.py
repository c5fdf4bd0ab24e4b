"""dev tool: issue rate of tcgen05.mma kind::tf32 (M 128, K 8) against N, accumulator rotation and the A operand's home"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyimsegm_b200 import _lib
torch = _lib.require_cuda()
lib = _lib.lib()
reps = 4000
print('cycles per instruction (1 CTA | 148 CTAs); ideal = N / 2 at 4096 tf32 FLOP/clk/SM')
print('%5s %22s %22s %22s %22s' % ('N', 'A smem, 1 acc', 'A smem, rotate', 'A tmem, 1 acc', 'A tmem, rotate'))
for N in (16, 32, 48, 64, 80, 96, 128, 160, 192, 240, 256):
    row = []
    for mode in (0, 1, 2, 3):
        cell = []
        for ctas in (1, 148):
            cyc = torch.zeros(ctas, dtype=torch.int64, device='cuda')
            for _ in range(2):
                _lib.check(lib.isb_umma_rate(N, reps, mode, ctas, _lib.ptr(cyc), _lib.stream_ptr()))
            torch.cuda.synchronize()
            cell.append(float(cyc.max()) / reps)
        row.append('%8.1f | %8.1f' % tuple(cell))
    print('%5d %22s %22s %22s %22s' % ((N,) + tuple(row)))
out = torch.zeros(4, dtype=torch.float64, device='cuda')
for _ in range(2):
    _lib.check(lib.isb_fp64_latency(8192, _lib.ptr(out), _lib.stream_ptr()))
torch.cuda.synchronize()
o = out.cpu().numpy()
print('dependent FP64 latency (clocks): add %.1f  mul %.1f  fma %.1f' % (o[0], o[1], o[2]))
