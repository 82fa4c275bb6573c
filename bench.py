#!/usr/bin/env python
"""Benchmark of the Diff-Reg denoising hot path on B200 (BASELINE.json metric: denoising steps/sec at
N=M=4096, d=256; Sinkhorn HBM GB/s vs peak).

    python bench.py --gpus N --steps K --warmup W            # this framework (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on the host

One "step" is one reverse-diffusion step of one 4DMatch-shaped sample (BASELINE.json configs[2]: N=M=4096,
d=256, Sinkhorn iters 3, eta=1 with noise), transformer excluded (features held fixed):
    mask + Sinkhorn(x_t) + exp -> top-K + SoftProcrustes + warp -> projection + similarity GEMM
    -> mask + Sinkhorn(sim) + exp -> mutual-NN matches at thr 0.2 -> DDIM update with fresh N(0,1) noise.
Each GPU runs its own independent sample (weak scaling, no collective on the data path).

Printed JSON line (rank 0): value = steps/s with inputs resident in HBM (CUDA-graph replay of the 20 step
graphs); e2e = the same step driven through the public modules with HOST (pinned) inputs copied in and the
step's results (pose, match count, matches) copied out every step; roofline = the fused Sinkhorn/DDIM call
timed with CUDA events inside an eager pass over the same steps; cpu_baseline = the oracle port on the host.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_PTS = 4096
FEAT_DIM = 256
SAMPLER_STEPS = 20
SKH_ITERS = 3
METRIC = "denoising steps/sec at N=M=4096,d=256"
UNIT = "steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_PTS, help="N = M (default: the headline 4096)")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="3xtf32", choices=["3xtf32", "tf32"])
    return ap.parse_args()


def workload_config(n, extra=None):
    cfg = {"workload": f"4DMatch diffusion sampler (BASELINE.json configs[2]): B=1 per GPU, N=M={n}, d={FEAT_DIM}, "
                       f"sampler steps={SAMPLER_STEPS}, sinkhorn iters={SKH_ITERS}, eta=1 with noise, thr=0.2",
           "N": n, "M": n, "d": FEAT_DIM, "sampler_steps": SAMPLER_STEPS, "skh_iters": SKH_ITERS}
    if extra:
        cfg.update(extra)
    return cfg


# --------------------------------------------------------------------------------------------
# synthetic inputs: unit-normal features, nn.Linear-style weight, rigidly related points
# --------------------------------------------------------------------------------------------
def make_inputs(seed, n, c):
    import torch
    g = torch.Generator().manual_seed(seed)
    src_feats = torch.randn(1, n, c, generator=g)
    tgt_feats = torch.randn(1, n, c, generator=g)
    bound = 1.0 / math.sqrt(c)
    W = (torch.rand(c, c, generator=g) * 2 - 1) * bound
    s_pcd = torch.randn(1, n, 3, generator=g)
    q = torch.randn(4, generator=g)
    q = q / q.norm()
    w, x, y, z = q.tolist()
    R = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    t = torch.randn(3, generator=g)
    perm = torch.randint(0, n, (n,), generator=g)
    t_pcd = (s_pcd[0, perm] @ R.t() + t + 0.01 * torch.randn(n, 3, generator=g))[None]
    ones = torch.ones(1, n, dtype=torch.bool)
    x_T = torch.randn(1, n, n, generator=g)
    return dict(src_feats=src_feats, tgt_feats=tgt_feats, W=W, s_pcd=s_pcd, t_pcd=t_pcd, src_mask=ones, tgt_mask=ones.clone(),
                x_T=x_T)


MATCH_CFG = dict(match_type="sinkhorn", confidence_threshold=0.2, feature_dim=FEAT_DIM, entangled=True, dsmax_temperature=0.1,
                 skh_init_bin_score=1.0, skh_iters=SKH_ITERS, skh_prefilter=False)


# --------------------------------------------------------------------------------------------
# clocks during the timed region
# --------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU leg: the oracle port (restatement of the reference's PyTorch code) on the host cores
# --------------------------------------------------------------------------------------------
def cpu_step_runner(n, seed=3000):
    """Returns (run_one_step, cores).  The oracle is the checker / CPU baseline only (oracle/ header)."""
    import torch
    from oracle import diffreg_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    inp = make_inputs(seed, n, FEAT_DIM)
    p = O.MatchingParams(src_proj_weight=inp["W"], bin_score=torch.tensor(1.0), skh_iters=SKH_ITERS)
    ac = O.alphas_cumprod()
    pairs = O.time_pairs(SAMPLER_STEPS)
    state = {"x": inp["x_T"].clone(), "k": 0}
    g = torch.Generator().manual_seed(seed + 1)

    def one_step():
        with torch.no_grad():
            t, tn = pairs[state["k"] % SAMPLER_STEPS]
            x = state["x"]
            O.noisy_matching_to_pose(x, p.bin_score, p.skh_iters, inp["s_pcd"], inp["t_pcd"], inp["src_mask"], inp["tgt_mask"],
                                     1.0, 40.0)
            sim, *_ = O.similarity(p, inp["src_feats"], inp["tgt_feats"])
            x0 = O.confidence_from_similarity(p, sim, inp["src_mask"], inp["tgt_mask"])
            O.get_match(x0, p.confidence_threshold)
            noise = torch.randn(x.shape, generator=g)
            state["x"] = O.ddim_update(x, x0, ac, t, tn, noise).float()
            state["k"] += 1

    return one_step, torch.get_num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    one_step, cores = cpu_step_runner(args.n)
    # bounded sample: every step is a full denoising step of the workload; at most ~2.5 min of timed CPU work
    # (about 3 s per step at the headline shape), so K large only raises the cap, not the run time
    for _ in range(min(args.warmup, 2)):
        one_step()
    t0 = time.perf_counter()
    done = 0
    while done < args.steps and (done < 3 or time.perf_counter() - t0 < 150.0):
        one_step()
        done += 1
    dt = time.perf_counter() - t0
    value = done / dt
    sample = (f"{done} full denoising steps (of K={args.steps} requested; capped at 150 s) at N=M={args.n} on {cores} host threads "
              f"(torch CPU, fp32; oracle port of the reference's PyTorch path)")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / done, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference", "config": workload_config(args.n),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    _emit(line)


# --------------------------------------------------------------------------------------------
# this framework
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from types import SimpleNamespace
    import diffreg_b200
    from diffreg_b200 import ops, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (diffreg_b200 has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n, c = args.n, FEAT_DIM
    host = make_inputs(3000 + rank, n, c)
    pinned = {k: host[k].pin_memory() for k in ("src_feats", "tgt_feats", "s_pcd", "t_pcd", "src_mask", "tgt_mask")}
    d = {k: v.to(dev) for k, v in host.items()}
    head = diffreg_b200.Matching(MATCH_CFG, precision=args.precision).to(dev).eval()
    with torch.no_grad():
        head.src_proj.weight.copy_(d["W"])
    proc = diffreg_b200.SoftProcrustesLayer(SimpleNamespace(sample_rate=1.0, max_condition_num=40.0))
    smp = diffreg_b200.DenoisingSampler("4d", head, proc, SAMPLER_STEPS, noise_seed=1234 + rank)
    feats = [d[k] for k in ("src_feats", "tgt_feats", "s_pcd", "t_pcd", "src_mask", "tgt_mask")]
    bufs = [d["x_T"].clone(), torch.empty_like(d["x_T"])]
    counter = torch.zeros(1, dtype=torch.int64, device=dev)

    def eager_step(i):
        k = i % SAMPLER_STEPS
        return smp.step(k, bufs[i % 2], None, *feats, x_out=bufs[(i + 1) % 2], noise_counter=counter)

    # ---- warm-up (eager) and graph capture of the 20 distinct steps
    for i in range(max(args.warmup, 3)):
        eager_step(i)
    torch.cuda.synchronize()
    graphs = None
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                eager_step(0)
                eager_step(1)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graphs = []
            pool = None
            for k in range(SAMPLER_STEPS):
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, pool=pool):
                    eager_step(k)
                pool = gph.pool()
                graphs.append(gph)
            for gph in graphs:        # first launch = upload of the executable graph to the device: setup, not a step
                gph.replay()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001 -- report and fall back to eager timing
            print(f"[bench] CUDA graph capture failed ({e}); timing eager launches", file=sys.stderr)
            graphs = None
    bufs[0].copy_(d["x_T"])

    def run_step(i):
        if graphs is not None:
            graphs[i % SAMPLER_STEPS].replay()
        else:
            eager_step(i)

    # kernels per step (counted once, eagerly)
    torch.cuda.synchronize()
    c0 = diffreg_b200.launch_count()
    eager_step(0)
    torch.cuda.synchronize()
    launches_per_step = diffreg_b200.launch_count() - c0
    bufs[0].copy_(d["x_T"])

    # ---- timed region 1: inputs resident in HBM
    for i in range(args.warmup):
        run_step(i)
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        run_step(args.warmup + i)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    clock_info = clocks.stop() if rank == 0 else None
    if dist is not None:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = world * args.steps / (ms * 1e-3)

    # ---- timed region 2 (e2e): host inputs in, step results out, every step, through the package's host front end
    #      (HostStepPipeline: pinned host inputs -> device on a copy stream overlapping the previous step, the step as a
    #      CUDA-graph replay of DenoisingSampler.step, pose / condition / match count / matches back to pinned host
    #      memory, one stream synchronisation per step because the caller consumes the result on the host)
    pipe = diffreg_b200.HostStepPipeline(smp, n, n, c, dev, use_graphs=not args.no_graph)
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes

    def e2e_run(first, count):
        # the host runs one step ahead: step i + 1 is enqueued (and step i + 2's input copy behind it) before step i's
        # results are awaited, so the GPU never idles while the host wakes up and reads a result
        end = first + count
        pipe.prefetch(first)
        pipe.launch(first)
        pipe.prefetch(first + 1)
        last = None
        for i in range(first, end):
            if i + 1 < end:
                pipe.launch(i + 1)
                pipe.prefetch(i + 2)
            last = int(pipe.finish(i)["count"][0])
        return last

    for slot in range(2):            # the step's inputs, written into the pipeline's pinned staging views (what a loader does)
        st_ = pipe.staging(slot)
        for k_ in st_:
            st_[k_].copy_(host[k_])
    pipe.reset(d["x_T"])
    w2 = max(args.warmup, 3)
    e2e_run(0, w2)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    e2e_run(w2, args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_launches = launches_per_step * args.steps      # graph nodes replayed (counted once, eagerly, above)
    barrier()
    if dist is not None:
        tt = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = world * args.steps / e2e_s

    # ---- roofline leg: eager pass over the same steps with the library's per-kernel CUDA-event hooks
    roofline = None
    kernel_ms = None
    if rank == 0:
        bufs[0].copy_(d["x_T"])
        for i in range(3):
            eager_step(i)
        torch.cuda.synchronize()
        _lib.profile_enable(True)
        ev = []
        for i in range(args.steps):
            eager_step(3 + i)
        torch.cuda.synchronize()
        prof = _lib.profile_read()
        _lib.profile_enable(False)
        kernel_ms = {k: round(v[0] / max(args.steps, 1), 5) for k, v in prof.items()}
        # dominant kernels: the fused Sinkhorn call = skh_persist_kernel (all iterations: one read of the matrix per
        # iteration, row and column log-sum-exp) + skh_final_kernel (exp / DDIM update).  Algorithmic bytes per call
        # are SURVEY.md section 8d's (2I+2) * E with E = 4 (N+1)(M+1).
        E = 4.0 * (n + 1) * (n + 1)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        it_ms, it_n = prof["skh_iter"]
        persistent = prof["skh_col"][1] == 0          # one launch runs all iterations
        skh_ms = sum(prof[k][0] for k in ("skh_prep", "skh_iter", "skh_col", "skh_final"))
        calls = 2 * args.steps
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if persistent and os.path.exists(tpath):      # dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full)
            tj = json.load(open(tpath)).get("skh_persist2_kernel")
            if tj:
                traffic = float(tj["dram_bytes_read"] + tj["dram_bytes_write"])
        if it_n and calls:
            # Dominant kernel: skh_persist2_kernel = ALL I iterations of one log_optimal_transport call in one launch
            # (row and column log-sum-exp of every iteration).  Algorithmic bytes per launch are SURVEY.md section 8d's
            # 2*I*E (I row-LSE reads + I column-LSE reads of the padded matrix, E = 4 (N+1)(M+1)); the final exp / DDIM
            # pass (the other 2E of the (2I+2)E call) is skh_final_tile_kernel, reported in "sinkhorn_call".
            # DRAM traffic is far below the algorithmic bytes because the row and column pass of an iteration share
            # one read and iterations 2..I hit the L2-resident matrix.
            it_s = it_ms * 1e-3 / it_n
            alg_it = ((2 * SKH_ITERS) if persistent else 2) * E
            per_call_s = skh_ms * 1e-3 / calls
            alg_call = (2 * SKH_ITERS + 2) * E
            roofline = {"bound": "hbm",
                        "kernel": "skh_persist2_kernel (register-slab persistent log-domain Sinkhorn, I=3 iterations per launch)"
                                  if persistent else "skh_iter*_kernel (one Sinkhorn iteration)",
                        "achieved": alg_it / it_s / 1e9, "peak": peak, "unit": "GB/s", "frac": alg_it / it_s / 1e9 / peak,
                        "traffic": traffic, "algorithmic_bytes_per_launch": alg_it, "us_per_launch": it_s * 1e6,
                        "launches": it_n, "peak_source": peak_src,
                        "sinkhorn_call": {"kernels": "skh_persist2_kernel + skh_final_tile_kernel (exp / DDIM + noise + arg-max pass)",
                                          "algorithmic_bytes": alg_call, "us_per_call": per_call_s * 1e6,
                                          "achieved": alg_call / per_call_s / 1e9, "frac": alg_call / per_call_s / 1e9 / peak}}

    # ---- CPU baseline (rank 0, single-GPU runs only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        one_step, cores = cpu_step_runner(n)
        one_step()
        t0 = time.perf_counter()
        reps = 0
        while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 12):
            one_step()
            reps += 1
        dt = time.perf_counter() - t0
        cpu = {"value": reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{reps} full denoising steps at N=M={n} (oracle port of the reference's PyTorch path, torch CPU fp32, "
                         f"after 1 warm-up step)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (tf32x3 tensor-core GEMM, fp32 accumulate)" if args.precision == "3xtf32" else "tf32",
                "data": "synthetic",
                "config": workload_config(n, {"timing": "cuda_graph_replay" if graphs is not None else "eager",
                                              "l2": "no explicit flush: each step streams five distinct 64 MiB fp32 matrices "
                                                    "(x_t, conf_d, sim, x0, x_next) > 126 MB L2",
                                              "precision": args.precision, "noise": "in-kernel Philox4x32-7"}),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(e2e_launches * world), "launches_per_step": int(launches_per_step),
                "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu, "kernel_ms_per_step": kernel_ms}
        _emit(line)
    if dist is not None:
        dist.destroy_process_group()


_JSON_FD = None


def _emit(line):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse_args()
    # Libraries write to stdout behind our back (NCCL prints its version line there on communicator creation): keep the
    # original stdout for the JSON line only and send everything else to stderr.
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
