"""Host-buffer front end of the denoising step: what a caller that keeps its data on the HOST uses.

Per step the caller's (pinned) inputs are copied into one of two device input sets on a copy stream -- the
copy of step i+1 overlaps the kernels of step i -- the step itself runs as ONE CUDA-graph replay of
DenoisingSampler.step (12 kernels, no host round trip inside), and the step's results (gated pose, condition
number, match count, matches) are copied back into pinned host buffers before `finish` returns.

    pipe = HostStepPipeline(sampler, n, m, c, device)
    pipe.reset(x_T)
    pipe.prefetch(0, inputs)                   # inputs: dict of pinned host tensors (src_feats, tgt_feats, s_pcd, ...)
    for i in range(steps):
        pipe.launch(i)                         # graph replay + result copies on the compute stream
        pipe.prefetch(i + 1, inputs)           # overlaps step i
        out = pipe.finish(i)                   # waits for step i only; pinned host tensors, valid until launch(i + 2)

(launch(i + 1) may also be issued before finish(i); on the B200 boxes measured here that was slower, the input copy
then competes with the kernels all the time -- tools/perf_e2e.py.)

The step index selects the DDIM time pair (i mod sampler.steps); consecutive steps alternate between the two
input sets and between the two state buffers, which is why sampler.steps must be even when graphs are used.
"""
import torch

from . import ops

_INPUT_KEYS = ("src_feats", "tgt_feats", "s_pcd", "t_pcd", "src_mask", "tgt_mask")


class HostStepPipeline:
    def __init__(self, sampler, n, m, c, device, use_graphs=True):
        if use_graphs and sampler.steps % 2:
            raise ValueError("HostStepPipeline: graph mode needs an even number of sampler steps")
        self.smp = sampler
        self.dev = torch.device(device)
        self.n, self.m, self.c = n, m, c
        shapes = {"src_feats": ((1, n, c), torch.float32), "tgt_feats": ((1, m, c), torch.float32),
                  "s_pcd": ((1, n, 3), torch.float32), "t_pcd": ((1, m, 3), torch.float32),
                  "src_mask": ((1, n), torch.bool), "tgt_mask": ((1, m), torch.bool)}
        self.sets = [{k: torch.zeros(s, dtype=dt, device=self.dev) for k, (s, dt) in shapes.items()} for _ in range(2)]
        for s in self.sets:       # a valid problem for the capture / warm-up runs
            s["src_mask"].fill_(True)
            s["tgt_mask"].fill_(True)
        self.x = [torch.zeros(1, n, m, device=self.dev), torch.zeros(1, n, m, device=self.dev)]
        # 3d flavour: the per-step x - x.min() (Diff-Reg-3dmatch pipeline.py:239) as a double-buffered device scalar:
        # step i reads shift[i % 2] and leaves min(x_next) in shift[(i + 1) % 2]
        self.shift = [torch.zeros(1, device=self.dev), torch.zeros(1, device=self.dev)] if sampler.flavour == "3d" else None
        self.counter = torch.zeros(1, dtype=torch.int64, device=self.dev)
        cap = min(n, m)
        self.host_out = [{"R": torch.empty(1, 3, 3).pin_memory(), "t": torch.empty(1, 3, 1).pin_memory(),
                          "condition": torch.empty(1, dtype=torch.float64).pin_memory(),
                          "count": torch.empty(1, dtype=torch.int32).pin_memory(),
                          "index": torch.empty(cap, 3, dtype=torch.int64).pin_memory(),
                          "mconf": torch.empty(cap).pin_memory()} for _ in range(2)]
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.sets[0].values())
        self.d2h_bytes = sum(v.numel() * v.element_size() for v in self.host_out[0].values())
        self.compute = torch.cuda.Stream(device=self.dev)
        self.copy = torch.cuda.Stream(device=self.dev)
        self.ev_in = [torch.cuda.Event(), torch.cuda.Event()]
        self.ev_free = [torch.cuda.Event(), torch.cuda.Event()]
        self.ev_out = [torch.cuda.Event(), torch.cuda.Event()]
        self.graphs = None
        self.aux = [None] * sampler.steps
        if use_graphs:
            self._capture()

    # ------------------------------------------------------------------------------------------
    def _eager(self, i):
        k = i % self.smp.steps
        s = self.sets[i % 2]
        shift = self.shift[i % 2] if self.shift is not None else None
        _, _, aux = self.smp.step(k, self.x[i % 2], shift, *[s[key] for key in _INPUT_KEYS], x_out=self.x[(i + 1) % 2],
                                  noise_counter=self.counter, x_min_out=self.shift[(i + 1) % 2] if self.shift is not None else None)
        return aux

    def _capture(self):
        with torch.cuda.stream(self.compute):
            self.x[0].normal_()
            for i in range(2):                 # warm-up: workspaces, kernel attributes
                self._eager(i)
        self.compute.synchronize()
        graphs, pool = [], None
        for k in range(self.smp.steps):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, stream=self.compute):
                self.aux[k] = self._eager(k)
            pool = g.pool()
            graphs.append(g)
        self.graphs = graphs
        self.compute.synchronize()

    # ------------------------------------------------------------------------------------------
    def reset(self, x_T):
        """Start a new sample from the noise state x_T (device or host tensor)."""
        with torch.cuda.stream(self.compute):
            self.x[0].copy_(x_T, non_blocking=True)
            if self.shift is not None:
                self.shift[0].copy_(ops.min_value(self.x[0]))
        self.compute.synchronize()

    def prefetch(self, i, inputs):
        """Enqueue the host -> device copy of step i's inputs (dict of pinned host tensors) on the copy stream."""
        s = self.sets[i % 2]
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(self.ev_free[i % 2])      # step i-2 (the last reader of this set) has run
            for key in _INPUT_KEYS:
                s[key].copy_(inputs[key], non_blocking=True)
            self.ev_in[i % 2].record(self.copy)

    def launch(self, i):
        """Run step i on the compute stream and enqueue the copy of its results to the host (all asynchronous)."""
        with torch.cuda.stream(self.compute):
            self.compute.wait_event(self.ev_in[i % 2])
            if self.graphs is not None:
                self.graphs[i % self.smp.steps].replay()
                aux = self.aux[i % self.smp.steps]
            else:
                aux = self._eager(i)
            self.ev_free[i % 2].record(self.compute)
            index, mconf, _, count = aux["match"]
            ho = self.host_out[i % 2]
            ho["R"].copy_(aux["pose"]["R_forwd"], non_blocking=True)
            ho["t"].copy_(aux["pose"]["t_forwd"], non_blocking=True)
            ho["condition"].copy_(aux["pose"]["condition"], non_blocking=True)
            ho["count"].copy_(count, non_blocking=True)
            ho["index"].copy_(index, non_blocking=True)
            ho["mconf"].copy_(mconf, non_blocking=True)
            self.ev_out[i % 2].record(self.compute)

    def finish(self, i):
        """Wait for step i's results; returns a dict of pinned host tensors (valid until launch(i + 2))."""
        self.ev_out[i % 2].synchronize()
        return self.host_out[i % 2]

    def state(self, i):
        """The device state x after step i-1 (input of step i)."""
        return self.x[i % 2]
