"""Phase timeline (globaltimer, ns) of procr_select_kernel and procr_moments_kernel inside the bench step (tuning only)."""
import sys, os, ctypes
os.environ["DRG_PROCR_TIMES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
import diffreg_b200, bench
lib = diffreg_b200.load_library()
dev = torch.device("cuda", 0)
n, c = 4096, 256
host = bench.make_inputs(3000, n, c)
d = {k: v.to(dev) for k, v in host.items()}
head = diffreg_b200.Matching(bench.MATCH_CFG).to(dev).eval()
with torch.no_grad():
    head.src_proj.weight.copy_(d["W"])
proc = diffreg_b200.SoftProcrustesLayer(SimpleNamespace(sample_rate=1.0, max_condition_num=40.0))
smp = diffreg_b200.DenoisingSampler("4d", head, proc, 20, noise_seed=1234)
feats = [d[k] for k in ("src_feats", "tgt_feats", "s_pcd", "t_pcd", "src_mask", "tgt_mask")]
bufs = [d["x_T"].clone(), torch.empty_like(d["x_T"])]
counter = torch.zeros(1, dtype=torch.int64, device=dev)
buf = (ctypes.c_longlong * 64)()
lib.drg_debug_read_procr_times.argtypes = [ctypes.c_void_p, ctypes.c_int]
for i in range(8):
    smp.step(i % 20, bufs[i % 2], None, *feats, x_out=bufs[(i + 1) % 2], noise_counter=counter)
    torch.cuda.synchronize()
    assert lib.drg_debug_read_procr_times(buf, 64) == 0
    t = list(buf)
    if i >= 4:
        print(f"step {i}: select: hist+walk {t[31]-t[30]} pass {t[32]-t[31]} rank {t[33]-t[32]} | select end -> moments start {t[40]-t[33]} | "
              f"moments CTA0: passA {t[41]-t[40]} passB {t[42]-t[41]} ticket {t[43]-t[42]} | last CTA: start->{t[44]-t[40]} combine {t[45]-t[44]} "
              f"kabsch {t[46]-t[45]} warp+end {t[47]-t[46]} | total moments {t[47]-t[40]} ns", flush=True)
