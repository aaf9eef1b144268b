"""Times the standalone device LU (lls_lu_factor + lls_lu_solve) on a synthetic diagonally dominant matrix.
usage: python scripts/lu_bench.py [n] [reps]"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from machupx_b200 import _native, layout as L  # noqa: E402
from machupx_b200.device import DeviceScene  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ld = L.padded_ld(n)
tables = {"n": n, "ld": ld, "n_aircraft": 1, "n_segments": 1, "flags": L.FLAG_DEFAULT,
          "tab": np.zeros((L.NF, ld)), "itab": np.zeros((L.NI, ld), dtype=np.int32), "airfoils": np.zeros((1, L.AF_STRIDE))}
dev = DeviceScene(tables)
g = torch.Generator(device="cuda").manual_seed(0)
# off-diagonal noise scaled by 1/sqrt(n): its spectral radius stays ~2 at every n, well inside the diagonal of 5..6, so the unpivoted
# factorisation is well conditioned at every size (an unscaled 0.05 randn outgrows the diagonal near n = 16000 and the residual
# of round 1's N = 16000 rows, 9.4e-9, measured the matrix, not the kernel)
A = torch.randn((ld, ld), dtype=torch.float64, device="cuda", generator=g) * (1.0 / n ** 0.5)
A += torch.diag(5.0 + torch.rand(ld, dtype=torch.float64, device="cuda", generator=g))
A[n:, :] = 0
A[:, n:] = 0
A[n:, n:] = torch.eye(ld - n, dtype=torch.float64, device="cuda")
b = torch.randn(ld, dtype=torch.float64, device="cuda", generator=g)
b[n:] = 0
lib = dev.lib
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for r in range(reps):
    M = A.clone()
    rhs = b.clone()
    torch.cuda.synchronize()
    e0.record()
    _native.check(lib.lls_lu_factor(dev.handle, M.data_ptr(), st))
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
_native.check(lib.lls_lu_solve(dev.handle, M.data_ptr(), rhs.data_ptr(), st))
torch.cuda.synchronize()
x = rhs[:n]
res = (A[:n, :n] @ x - b[:n]).abs().max().item() / b[:n].abs().max().item()
flops = 2.0 / 3.0 * n ** 3
print(json.dumps({"n": n, "ld": ld, "lu_ms": best, "tflops": flops / best * 1e-9, "rel_residual": res}))
