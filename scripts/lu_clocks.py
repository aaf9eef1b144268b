"""Phase clocks of the LU diag CTA (debug build liblls_clk.so, -DLLS_LU_CLOCKS).  usage: python scripts/lu_clocks.py [n]"""
import ctypes, os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from machupx_b200 import _native, layout as L
_native.LIB_PATH = os.path.join(os.path.dirname(_native.LIB_PATH), "liblls_clk.so")
from machupx_b200.device import DeviceScene
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ld = L.padded_ld(n)
tables = {"n": n, "ld": ld, "n_aircraft": 1, "n_segments": 1, "flags": L.FLAG_DEFAULT, "tab": np.zeros((L.NF, ld)),
          "itab": np.zeros((L.NI, ld), dtype=np.int32), "airfoils": np.zeros((1, L.AF_STRIDE))}
dev = DeviceScene(tables)
A = torch.randn((ld, ld), dtype=torch.float64, device="cuda") * 0.05 + torch.diag(5.0 + torch.rand(ld, dtype=torch.float64, device="cuda"))
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for rep in range(3):
    M = A.clone()
    _native.check(dev.lib.lls_lu_factor(dev.handle, M.data_ptr(), st))
    torch.cuda.synchronize()
out = (ctypes.c_longlong * 16)()
dev.lib.lls_debug_lu_clocks(out)
c = list(out)
names = ["start", "loaded", "mma", "pre-factor", "factored", "published", "pre-solve", "solved", "end"]
print(json.dumps({names[i + 1] + "-" + names[i]: c[i + 1] - c[i] for i in range(8)}))
