import sys, os, ctypes
os.environ["DRG_PROCR_TIMES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffreg_b200
from diffreg_b200 import ops
import bench
lib = diffreg_b200.load_library()
n = 4096
host = bench.make_inputs(3000, n, 256)
d = {k: v.cuda() for k, v in host.items()}
alpha = torch.tensor(1.0, device="cuda")
conf = ops.sinkhorn(d["x_T"], alpha, 3, d["src_mask"], d["tgt_mask"], out_mode="conf", apply_mask=True)
for _ in range(3):
    o = ops.soft_procrustes(conf, d["s_pcd"], d["t_pcd"], d["src_mask"], d["tgt_mask"], 1.0, 40.0, want_warped=True)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
lib.drg_debug_read_procr_times.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.drg_debug_read_procr_times(buf, 64) == 0
t = list(buf)
print("emit: compact", t[7] - t[2], "dense", t[3] - t[7])
print("solve: stage", t[1] - t[0], "select", t[2] - t[1], "emit", t[3] - t[2], "block sums", t[4] - t[3], "kabsch(thread0)", t[5] - t[4], "warp+end", t[6] - t[5], "total", t[6] - t[0], "n_cand", t[10])
print("threshold: counts", t[21] - t[20], "sample", t[22] - t[21], "select", t[23] - t[22], "total", t[23] - t[20])
