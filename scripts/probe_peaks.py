import ctypes, json, torch
torch.cuda.init()
lib = ctypes.CDLL("machupx_b200/liblls.so")
out = (ctypes.c_double*4)()
rc = lib.lls_probe_peaks(out, None)
print(json.dumps({"rc": rc, "dfma_tflops": out[0], "dmma884_tflops": out[1], "dmma16816_tflops": out[2], "copy_gbs": out[3]}))
