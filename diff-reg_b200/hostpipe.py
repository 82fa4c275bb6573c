"""Host-buffer front end of the denoising step: what a caller that keeps its data on the HOST uses.

Per step the caller's inputs travel host -> device as ONE copy: the six input tensors of a step live back to back in
one pinned staging buffer (`staging(i)` hands out the views the caller fills: a data loader writes its features, points
and masks straight into them) and in one device buffer of the same layout, so `prefetch(i)` is a single 8.5 MB
cudaMemcpyAsync on the copy stream (~50 GB/s on PCIe Gen5; six separate copies of 4 MB / 48 KB / 4 KB ran at ~20 GB/s in
total) that overlaps the kernels of step i-1.  The step itself runs as ONE CUDA-graph replay of DenoisingSampler.step
(no host round trip inside), and the step's results (gated pose, condition number, match count, matches) are copied back
into pinned host buffers before `finish` returns.

    pipe = HostStepPipeline(sampler, n, m, c, device)
    pipe.reset(x_T)
    pipe.staging(0)["src_feats"].copy_(...)    # fill step 0's pinned views (or: pipe.prefetch(0, dict_of_host_tensors))
    pipe.prefetch(0)
    for i in range(steps):
        pipe.launch(i)                         # graph replay + result copies on the compute stream
        pipe.prefetch(i + 1)                   # one H2D copy, overlaps step i (staging(i + 1) filled by the caller)
        out = pipe.finish(i)                   # waits for step i only; pinned host tensors, valid until launch(i + 3)

launch(i + 1) and launch(i + 2) may be issued before finish(i): the input sets and state buffers are double-buffered
(prefetch(i + 2) waits on the device for step i, the last reader of its set) and the result buffers are three deep, so
the host can run up to TWO steps ahead and a host hiccup of a whole step time does not idle the GPU.  Order per step:
prefetch(j) before launch(j).

The step index selects the DDIM time pair (i mod sampler.steps); consecutive steps alternate between the two
input sets and between the two state buffers, which is why sampler.steps must be even when graphs are used.
"""
import torch

from . import ops

_INPUT_KEYS = ("src_feats", "tgt_feats", "s_pcd", "t_pcd", "src_mask", "tgt_mask")


class HostStepPipeline:
    def __init__(self, sampler, n, m, c, device, use_graphs=True, results_stream=True):
        if use_graphs and sampler.steps % 2:
            raise ValueError("HostStepPipeline: graph mode needs an even number of sampler steps")
        self.smp = sampler
        self.dev = torch.device(device)
        self.n, self.m, self.c = n, m, c
        shapes = {"src_feats": ((1, n, c), torch.float32), "tgt_feats": ((1, m, c), torch.float32),
                  "s_pcd": ((1, n, 3), torch.float32), "t_pcd": ((1, m, 3), torch.float32),
                  "src_mask": ((1, n), torch.bool), "tgt_mask": ((1, m), torch.bool)}
        # one packed buffer per input set (device) and per staging slot (pinned host), sections 256-byte aligned
        offs, off = {}, 0
        for k, (shp, dt) in shapes.items():
            nbytes = int(torch.tensor([], dtype=dt).element_size())
            for d_ in shp:
                nbytes *= d_
            offs[k] = (off, nbytes)
            off = (off + nbytes + 255) // 256 * 256
        self.packed_bytes = off

        def views(buf):
            return {k: buf[o:o + nb].view(shapes[k][1]).view(shapes[k][0]) for k, (o, nb) in offs.items()}

        self.dev_packed = [torch.zeros(self.packed_bytes, dtype=torch.uint8, device=self.dev) for _ in range(2)]
        self.host_packed = [torch.zeros(self.packed_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.sets = [views(b) for b in self.dev_packed]
        self.host_sets = [views(b) for b in self.host_packed]
        for s in self.sets + self.host_sets:       # a valid problem for the capture / warm-up runs
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
                          "mconf": torch.empty(cap).pin_memory()} for _ in range(3)]
        self.h2d_bytes = self.packed_bytes        # what one prefetch moves (the six tensors + < 1.5 KB of alignment padding)
        self.d2h_bytes = sum(v.numel() * v.element_size() for v in self.host_out[0].values())
        self.compute = torch.cuda.Stream(device=self.dev)
        self.copy = torch.cuda.Stream(device=self.dev)        # host -> device
        self.out = torch.cuda.Stream(device=self.dev) if results_stream else self.compute   # device -> host (results)
        self.ev_in = [torch.cuda.Event(), torch.cuda.Event()]
        self.ev_free = [torch.cuda.Event(), torch.cuda.Event()]
        self.ev_out = [torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()]
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
        # a graph's first launch uploads it to the device (hundreds of microseconds): pay that here, not in the caller's loop
        with torch.cuda.stream(self.compute):
            for g in graphs:
                g.replay()
        self.compute.synchronize()

    # ------------------------------------------------------------------------------------------
    def reset(self, x_T):
        """Start a new sample from the noise state x_T (device or host tensor)."""
        with torch.cuda.stream(self.compute):
            self.x[0].copy_(x_T, non_blocking=True)
            if self.shift is not None:
                self.shift[0].copy_(ops.min_value(self.x[0]))
        self.compute.synchronize()

    def staging(self, i):
        """The pinned host views of step i's inputs (dict: src_feats, tgt_feats, s_pcd, t_pcd, src_mask, tgt_mask) inside
        the packed staging buffer of slot i % 2.  Fill them, then call prefetch(i).  A slot may be rewritten once
        prefetch(i)'s copy has been consumed, i.e. after finish(i) returned."""
        return self.host_sets[i % 2]

    def prefetch(self, i, inputs=None):
        """Enqueue the host -> device copy of step i's inputs on the copy stream.
        inputs=None: ONE copy of the packed staging buffer of slot i % 2 (filled through staging(i)).
        inputs=dict of host tensors: copied tensor by tensor (six copies; for callers that cannot write into staging)."""
        s = self.sets[i % 2]
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(self.ev_free[i % 2])      # step i-2 (the last reader of this set) has run
            if inputs is None:
                self.dev_packed[i % 2].copy_(self.host_packed[i % 2], non_blocking=True)
            else:
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
        # the result copies run on their own stream: the next step's kernels do not wait for them (every step graph owns its
        # output tensors, and in eager mode the tensors are fresh per call)
        with torch.cuda.stream(self.out):
            self.out.wait_event(self.ev_free[i % 2])
            index, mconf, _, count = aux["match"]
            ho = self.host_out[i % 3]
            ho["R"].copy_(aux["pose"]["R_forwd"], non_blocking=True)
            ho["t"].copy_(aux["pose"]["t_forwd"], non_blocking=True)
            ho["condition"].copy_(aux["pose"]["condition"], non_blocking=True)
            ho["count"].copy_(count, non_blocking=True)
            ho["index"].copy_(index, non_blocking=True)
            ho["mconf"].copy_(mconf, non_blocking=True)
            self.ev_out[i % 3].record(self.out)
            if self.graphs is None:
                for t_ in (aux["pose"]["R_forwd"], aux["pose"]["t_forwd"], aux["pose"]["condition"], count, index, mconf):
                    t_.record_stream(self.out)

    def finish(self, i):
        """Wait for step i's results; returns a dict of pinned host tensors (valid until launch(i + 3))."""
        self.ev_out[i % 3].synchronize()
        return self.host_out[i % 3]

    def state(self, i):
        """The device state x after step i-1 (input of step i)."""
        return self.x[i % 2]
