"""A/B of the host-pipeline loop orders (tuning only)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
import diffreg_b200, bench
n, c = 4096, 256
dev = torch.device("cuda", 0)
host = bench.make_inputs(3000, n, c)
pinned = {k: host[k].pin_memory() for k in ("src_feats", "tgt_feats", "s_pcd", "t_pcd", "src_mask", "tgt_mask")}
head = diffreg_b200.Matching(bench.MATCH_CFG).to(dev).eval()
with torch.no_grad():
    head.src_proj.weight.copy_(host["W"].to(dev))
proc = diffreg_b200.SoftProcrustesLayer(SimpleNamespace(sample_rate=1.0, max_condition_num=40.0))
smp = diffreg_b200.DenoisingSampler("4d", head, proc, 20, noise_seed=1)
pipe = diffreg_b200.HostStepPipeline(smp, n, n, c, dev)
pipe.reset(host["x_T"])
def run_ahead(first, count):
    end = first + count
    pipe.prefetch(first, pinned); pipe.launch(first); pipe.prefetch(first + 1, pinned)
    for i in range(first, end):
        if i + 1 < end:
            pipe.launch(i + 1); pipe.prefetch(i + 2, pinned)
        pipe.finish(i)
def run_plain(first, count):
    pipe.prefetch(first, pinned)
    for i in range(first, first + count):
        pipe.launch(i); pipe.prefetch(i + 1, pinned); pipe.finish(i)
def run_nocopy(first, count):
    for i in range(first, first + count):
        pipe.graphs[i % 20].replay()
    torch.cuda.synchronize()
# host time spent inside each call (perf_counter), per loop order
import collections
_acc = collections.defaultdict(float)
def _timed(name, fn):
    def w(*a, **k):
        t = time.perf_counter()
        r = fn(*a, **k)
        _acc[name] += time.perf_counter() - t
        return r
    return w
pipe.launch = _timed("launch", pipe.launch); pipe.prefetch = _timed("prefetch", pipe.prefetch); pipe.finish = _timed("finish", pipe.finish)
K = 200
pos = 0
for name, fn in (("plain", run_plain), ("ahead", run_ahead), ("plain", run_plain), ("ahead", run_ahead)):
    fn(pos, 20); pos += 20
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fn(pos, K); pos += K
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(name, round(K / dt, 1), "steps/s", round(1e6 * dt / K, 1), "us/step",
          {k: round(1e6 * v / (K + 20), 1) for k, v in _acc.items()}, "host us/step", flush=True)
    _acc.clear()
with torch.cuda.stream(pipe.compute):
    run_nocopy(0, 20)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    run_nocopy(0, K)
    dt = time.perf_counter() - t0
print("graphs only", round(K / dt, 1), "steps/s")
