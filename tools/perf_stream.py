"""Pure streaming-read micro-benchmarks (TMA bulk ring vs LDG.128) to calibrate what the Sinkhorn passes can reach."""
import sys, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffreg_b200
lib = diffreg_b200.load_library()
lib.drg_debug_stream.restype = ctypes.c_int
lib.drg_debug_stream.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
dev = "cuda"
out = torch.zeros(4, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def run(x, mode, stage_floats, nstage, grid, reps=10, flush_l2=True):
    st = torch.cuda.current_stream().cuda_stream
    ts = []
    for r in range(reps + 2):
        if flush_l2: flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.drg_debug_stream(x.data_ptr(), x.numel(), mode, stage_floats, nstage, grid, out.data_ptr(), st)
        assert rc == 0, lib.drg_last_error()
        e1.record(); torch.cuda.synchronize()
        if r >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
for n in (4096, 16384):
    x = torch.randn(n, n, device=dev)
    nbytes = x.numel() * 4
    for (sf, ns) in [(16384, 3), (8192, 6), (4096, 12), (2048, 24), (1024, 48)]:
        us = run(x, 0, sf, ns, 148)
        print(json.dumps(dict(n=n, mode="tma_ring", stage_kb=sf * 4 // 1024, nstage=ns, us=round(us, 1), GBps=round(nbytes / us / 1e3))), flush=True)
    for grid in (148, 296, 592, 1184):
        us = run(x, 1, 0, 0, grid)
        print(json.dumps(dict(n=n, mode="ldg128", grid=grid, us=round(us, 1), GBps=round(nbytes / us / 1e3))), flush=True)
    if n == 4096:
        us = run(x, 0, 16384, 3, 148, flush_l2=False)
        print(json.dumps(dict(n=n, mode="tma_ring_noflush", us=round(us, 1), GBps=round(nbytes / us / 1e3))), flush=True)
        us = run(x, 1, 0, 0, 592, flush_l2=False)
        print(json.dumps(dict(n=n, mode="ldg128_noflush", us=round(us, 1), GBps=round(nbytes / us / 1e3))), flush=True)
    del x
