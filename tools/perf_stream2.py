import sys, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffreg_b200
lib = diffreg_b200.load_library()
lib.drg_debug_stream.restype = ctypes.c_int
lib.drg_debug_stream.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
dev = "cuda"
out = torch.zeros(4, device=dev)
st = torch.cuda.current_stream().cuda_stream
for mb in (4, 16, 32, 43, 48, 64, 96):
    n = mb * 1024 * 1024 // 4
    x = torch.randn(n, device=dev)
    for mode, grid, sf, ns in ((1, 592, 0, 0), (0, 148, 4096, 12), (0, 148, 16384, 3)):
        for _ in range(3):
            lib.drg_debug_stream(x.data_ptr(), n, mode, sf, ns, grid, out.data_ptr(), st)
        torch.cuda.synchronize()
        reps = 20
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            lib.drg_debug_stream(x.data_ptr(), n, mode, sf, ns, grid, out.data_ptr(), st)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        print(json.dumps(dict(MB=mb, mode=["tma", "ldg"][mode], stage_kb=sf * 4 // 1024, us_back_to_back=round(us, 2), GBps=round(n * 4 / us / 1e3))), flush=True)
    del x
