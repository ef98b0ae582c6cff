import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import allhic_b200 as A
from allhic_b200 import synth
import torch
ids, clmf, _ = synth.materialize("config3")
clm = A.CLM(clmf, ids); eng = A.Engine(0); eng.load(clm)
n, P = 2000, 8000
rng = np.random.default_rng(0)
tours = np.stack([rng.permutation(n) for _ in range(P)]).astype(np.uint16)
pin_t = A.PinnedArray((P, n), np.uint16); pin_o = A.PinnedArray((P,), np.float64); pin_t.array[:] = tours
for ch in (1, 2, 4, 8, 16, 32):
    os.environ["AHX_EVAL_CHUNKS"] = str(ch)
    for _ in range(5): eng.eval_m(pin_t.array, out=pin_o.array)
    t0 = time.perf_counter()
    for _ in range(30): eng.eval_m(pin_t.array, out=pin_o.array)
    dt = (time.perf_counter() - t0) / 30
    print("chunks %2d: %.3f ms  %.2f M evals/s  %.1f GB/s" % (ch, dt * 1e3, P / dt / 1e6, pin_t.nbytes / dt / 1e9))
# raw pinned H2D ceiling
x = torch.empty(32_000_000, dtype=torch.uint8).pin_memory(); y = torch.empty_like(x, device="cuda")
for _ in range(3): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print("plain pinned H2D 32 MB: %.3f ms %.1f GB/s" % (dt * 1e3, 32e6 / dt / 1e9))
