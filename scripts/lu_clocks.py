"""Phase clocks of the LU chain per block step (debug build liblls_clk.so, -DLLS_LU_CLOCKS): the diag CTA and the first
L-panel CTA stamp clock64 at every stage.  usage: python scripts/lu_clocks.py [n]"""
import ctypes, os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from machupx_b200 import _native, layout as L
_native.LIB_PATH = os.path.join(os.path.dirname(_native.LIB_PATH), "liblls_clk.so")
from machupx_b200.device import DeviceScene
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
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
steps = min(ld // 64, 256)
out = (ctypes.c_longlong * (24 * steps))()
dev.lib.lls_debug_lu_clocks(out, steps)
c = np.array(list(out), dtype=np.int64).reshape(steps, 24)
rows = []
for k in range(steps):
    d, p = c[k, :9], c[k, 12:21]
    rows.append({"block": k, "rem": steps - k, "diag_load": int(d[1] - d[0]), "diag_mma": int(d[2] - d[1]), "factor": int(d[4] - d[3]),
                 "publish": int(d[5] - d[4]), "diag_total": int(d[8] - d[0]), "panel_wait+copy": int(p[6] - p[2]), "panel_solve": int(p[7] - p[6]),
                 "panel_total": int(p[8] - p[0])})
for r in rows[:: max(1, steps // 32)]:
    print(json.dumps(r))
