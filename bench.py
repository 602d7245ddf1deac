#!/usr/bin/env python
"""bench.py -- WKV-7 hot path, BASELINE.json config 2: RWKV-7 0.4B Spark layout, bf16,
batch 8, seq_len 4096 per GPU (H=16 heads of 64, L=24 layers).

A "step" is one pass of the hot path over one batch: the WKV-7 forward of all 24 layers followed
by the WKV-7 backward of all 24 layers on [8,4096,16,64] synthetic inputs.
metric = audio-tokens/s = B*T*n_gpus / step time.  Shards are independent (pure data parallel:
each rank owns its own batch of 8 sequences, no data-path collective) -> "scaling": "weak".

  python bench.py [--gpus N --steps K --warmup W]          our arm (device-resident `value`, `e2e`
                                                           through the public API with host buffers)
  python bench.py --impl reference ...                     the reference algorithm on the host cores
                                                           (oracle C port, OpenMP), bounded sample

Under torchrun (N > 1) every rank runs its shard; rank 0 prints the single JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, T, H, C, LAYERS = 8, 4096, 16, 64, 24
FWD_BYTES, BWD_BYTES = 7 * C * 2, 13 * C * 2          # algorithmic bytes per token-head
METRIC = "audio-tokens/sec (train fwd+bwd) RWKV-7 0.4B seq4096"
WORKLOAD = "configs[1]: RWKV-7 0.4B Spark-layout bf16, batch 8/GPU, seq_len 4096, WKV-7 fwd+bwd x 24 layers"


_CTL = {"mode": "single", "dir": None, "seq": 0, "rank": 0}
_T0 = time.time()


def log(msg):
    """One stderr line per leg: a hang can no longer erase the record of how far the run got."""
    sys.stderr.write("bench[%s +%.1fs]: %s\n" % (os.environ.get("RANK", "0"), time.time() - _T0, msg))
    sys.stderr.flush()


def dist_init(world):
    """Control plane of a multi-GPU run.  The path shards by batch with no data-path collective, so
    the only exchanges of the bench are its barriers and the max over ranks of two scalars: they go over gloo (CPU
    tensors).  The process group is created with NCCL registered for CUDA tensors, as a training job would have it
    (the ZeRO-2 engine's reduce-scatter / all-gather), but the NCCL communicator is only built on the first CUDA
    collective -- which this bench never issues -- so an 8-rank start does not pay (or hang in) NVLS / fabric set-up.
    If the process group cannot be created within two minutes (or RWKVTTS_BENCH_CONTROL=fs), the same two exchanges run
    over files in /tmp: the ranks of one torchrun launch share a node and a parent process."""
    if world <= 1:
        return
    import datetime
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if os.environ["MASTER_ADDR"] in ("127.0.0.1", "localhost"):
        os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")     # one node: do not depend on the hostname resolving
    _CTL["rank"] = int(os.environ.get("RANK", "0"))
    _CTL["dir"] = os.path.join("/tmp", "rwkvtts_bench_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid()))
    import torch
    if os.environ.get("RWKVTTS_BENCH_CONTROL") != "fs":
        tmo = datetime.timedelta(seconds=120)
        for backend in (("cpu:gloo,cuda:nccl",) if torch.cuda.is_available() else ()) + ("gloo",):
            try:
                dist.init_process_group(backend, timeout=tmo)
                dist.all_reduce(torch.zeros(1))               # proves the gloo ring before anything is timed
                _CTL["mode"] = "dist"
                return
            except Exception as e:                            # mixed registration unavailable / rendezvous failed
                sys.stderr.write("bench: process group %r failed (%r)\n" % (backend, e))
                if dist.is_initialized():
                    dist.destroy_process_group()
    os.makedirs(_CTL["dir"], exist_ok=True)
    _CTL["mode"] = "fs"


def dist_finish(world):
    if world > 1 and _CTL["mode"] == "dist":
        import torch.distributed as dist
        dist.destroy_process_group()


def _fs_exchange(x, world, timeout_s=900.0):
    """Every rank publishes one float for this sequence number and reads everybody's (atomic rename, polling)."""
    seq, _CTL["seq"] = _CTL["seq"], _CTL["seq"] + 1
    d, r = _CTL["dir"], _CTL["rank"]
    tmp = os.path.join(d, ".%d.%d.tmp" % (seq, r))
    with open(tmp, "w") as f:
        f.write(repr(float(x)))
    os.replace(tmp, os.path.join(d, "%d.%d" % (seq, r)))
    vals, t0 = {}, time.time()
    while len(vals) < world:
        for k in range(world):
            if k not in vals:
                try:
                    vals[k] = float(open(os.path.join(d, "%d.%d" % (seq, k))).read())
                except (OSError, ValueError):
                    pass
        if len(vals) < world:
            if time.time() - t0 > timeout_s:
                raise RuntimeError("bench control plane: rank(s) %s never reached exchange %d"
                                   % (sorted(set(range(world)) - set(vals)), seq))
            time.sleep(0.002)
    return [vals[k] for k in range(world)]


def dist_barrier(world):
    import torch
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if world > 1:
        if _CTL["mode"] == "fs":
            _fs_exchange(0.0, world)
            return
        import torch.distributed as dist
        dist.all_reduce(torch.zeros(1))          # CPU tensor -> gloo


def dist_max(x, world):
    import torch
    if world <= 1:
        return float(x)
    if _CTL["mode"] == "fs":
        return max(_fs_exchange(x, world))
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference(seconds_target=12.0):
    """The reference algorithm (C port of forward_kernel/backward_kernel) on the host cores, on a
    bounded sample: one layer-call of [b,4096,16,64]; tokens/s scaled to the 24-layer step."""
    import torch
    from oracle import c_oracle as CO
    from oracle import wkv7_oracle as O
    CO.build(ref=False)
    cores = CO.num_threads()
    b = min(B, max(1, -(-2 * cores // H)))      # at least 2 (b,h) units per thread
    x = O.make_inputs(b, T, H, seed=42)
    args = [x[n] for n in "wqkvab"]
    t0 = time.perf_counter()
    y, s, sa = CO.c_forward(*args)
    CO.c_backward(*args, x["dy"], s, sa)
    dt = time.perf_counter() - t0                       # warm-up + calibration
    reps = max(1, min(8, int(seconds_target / max(dt, 1e-3))))
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        y, s, sa = CO.c_forward(*args)
        CO.c_backward(*args, x["dy"], s, sa)
        best = min(best, time.perf_counter() - t0)
    value = b * T / (best * LAYERS)
    try:
        model = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        model = "unknown"
    return {"value": value, "unit": "tokens/s", "cores": cores, "kind": "port",
            "sample": f"1 of {LAYERS} layers, batch {b} of {B}, T={T}, H={H}: fwd+bwd best of {reps} "
                      f"= {best * 1e3:.0f} ms, scaled x{LAYERS} layers; C port of wkv7_cuda.cu "
                      f"forward_kernel/backward_kernel, OpenMP over (b,h); cpu: {model}"}


def decode_leg(dev, new_tokens=128):
    """AR decode, BASELINE config c4: RWKV-7 0.4B (D=1024, L=24, H=16, vocab 8193), 32 prompts of 163 positions, greedy,
    EOS suppressed, through RWKV7ForCausalLM.generate (CUDA-graph step).  tokens/s = 32 * steps / time after prefill."""
    import torch
    from rwkvfla.models.rwkv7 import RWKV7Config, RWKV7ForCausalLM
    DB, PROMPT = 32, 163
    torch.manual_seed(42)
    cfg = RWKV7Config(hidden_size=1024, num_hidden_layers=24, head_dim=64, vocab_size=8193, decay_low_rank_dim=64,
                      a_low_rank_dim=64, v_low_rank_dim=32, gate_low_rank_dim=128)
    m = RWKV7ForCausalLM(cfg)
    with torch.no_grad():
        for _, p in m.named_parameters():
            if p.abs().sum() == 0:
                p.copy_(torch.randn_like(p) * 0.02)
    m = m.to(dev).to(torch.bfloat16).eval()
    ids = torch.randint(0, 8192, (DB, PROMPT), device=dev)
    out = {}
    for mode, steps in (("cuda_graph", new_tokens), ("eager", 40)):
        kw = dict(input_ids=ids, do_sample=False, eos_token_id=None, use_cuda_graph=(mode == "cuda_graph"))
        m.generate(max_new_tokens=10, **kw)
        torch.cuda.synchronize()
        short = 16                     # both runs pay prefill (+ graph capture): the difference is pure decode steps
        t0 = time.perf_counter(); m.generate(max_new_tokens=short, **kw); torch.cuda.synchronize()
        t_short = time.perf_counter() - t0
        t0 = time.perf_counter(); seq = m.generate(max_new_tokens=steps, **kw); torch.cuda.synchronize()
        dt = time.perf_counter() - t0 - t_short
        out[mode] = {"tokens_per_s": DB * (steps - short) / dt, "ms_per_step": dt / (steps - short) * 1e3, "steps": steps,
                     "prefill_plus_setup_ms": (t_short - short * dt / (steps - short)) * 1e3}
        out[mode + "_ids"] = seq[:, PROMPT:PROMPT + 24]
    same = bool(torch.equal(out.pop("cuda_graph_ids"), out.pop("eager_ids")))
    out["greedy_ids_identical_graph_vs_eager"] = same
    out["config"] = "configs[3]: RWKV-7 0.4B random init, batch 32, prompt 163, greedy, EOS suppressed; steps timed as the difference of a long and a 16-token run"
    del m
    torch.cuda.empty_cache()
    return out


def model_step_leg(dev, steps=2):
    """Whole-model training step (forward + backward) at BASELINE config c2: RWKV-7 0.4B (D=1024, L=24, H=16, vocab 8193),
    batch 8 x 4096, bf16, random init, synthetic ids: tokens/s with the fused time-mix kernels and with the ATen elementwise
    chain around the same WKV kernels (what the kernels outside the recurrence buy end to end)."""
    import torch
    from rwkvtts_b200 import core
    from rwkvfla.models.rwkv7 import RWKV7Config, RWKV7ForCausalLM
    torch.manual_seed(42)
    cfg = RWKV7Config(hidden_size=1024, num_hidden_layers=24, head_dim=64, vocab_size=8193, decay_low_rank_dim=64,
                      a_low_rank_dim=64, v_low_rank_dim=32, gate_low_rank_dim=128, fuse_cross_entropy=True)
    m = RWKV7ForCausalLM(cfg)
    with torch.no_grad():
        for _, p in m.named_parameters():
            if p.abs().sum() == 0:
                p.copy_(torch.randn_like(p) * 0.02)
    m = m.to(dev).to(torch.bfloat16).train()
    ids = torch.randint(0, 8192, (B, T), device=dev)
    out = {}
    try:
        for name, flag in (("fused", True), ("aten", False)):
            core.FUSED = flag
            for _ in range(2):
                m.zero_grad(set_to_none=True)
                m(input_ids=ids, labels=ids).loss.backward()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                m.zero_grad(set_to_none=True)
                loss = m(input_ids=ids, labels=ids).loss
                loss.backward()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"ms_per_step": ms, "tokens_per_s": B * T / ms * 1e3, "loss": float(loss.detach())}
    finally:
        core.FUSED = True
    out["config"] = "configs[1] whole model: RWKV-7 0.4B, batch 8 x 4096, fwd+bwd, 1 GPU; fused = csrc/tmix_fused.cu, aten = ATen chain"
    del m
    torch.cuda.empty_cache()
    return out


def run_reference(args, rank):
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone and may use every host
    # core (the other ranks exit without work), so undo that before the OpenMP runtime of the C port starts
    if "LOCAL_RANK" in os.environ or os.environ.get("OMP_NUM_THREADS") == "1":
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity")
                                            else (os.cpu_count() or 1))
    cb = cpu_reference(seconds_target=20.0)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": B * T / cb["value"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 in / fp32 state",
            "data": "synthetic", "config": {"workload": WORKLOAD}, "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-model-step", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist_init(world)
    extras = world == 1          # the explanatory legs (other kernels, fused kernels, reference CUDA op) run at N = 1 only
    import rwkvtts_b200 as R
    from rwkvtts_b200.synth import make_inputs
    lib = R._lib.lib()

    x = make_inputs(B, T, H, seed=42 + rank)
    d = {n: t.to(dev) for n, t in x.items()}
    ins = [d[n] for n in "wqkvab"]
    y = torch.empty_like(d["v"])
    s = torch.empty(B, H, T // 16, C, C, dtype=torch.float32, device=dev)
    sa = torch.empty(B, T, H, C, dtype=torch.float32, device=dev)
    grads = [torch.empty_like(d["v"]) for _ in range(6)]

    def step(ev_mid=None):
        for _ in range(LAYERS):
            R.wkv7_forward_(*ins, y, s, sa)
        if ev_mid is not None:
            ev_mid.record()
        for _ in range(LAYERS):
            R.wkv7_backward_(*ins, d["dy"], s, sa, *grads)

    def barrier():
        dist_barrier(world)

    log("inputs resident; warm-up")
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    log("warm-up done; timed region")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.rwkvtts_kernel_launches()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        ev[i][0].record()
        step(ev[i][1])
        ev[i][2].record()
    barrier()
    log("timed region done")
    launches = lib.rwkvtts_kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0][0].elapsed_time(ev[-1][2])
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / (args.steps * LAYERS)
    bwd_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / (args.steps * LAYERS)
    total_ms = dist_max(total_ms, world)
    ms_per_step = total_ms / args.steps
    value = B * T * world / (ms_per_step * 1e-3)

    # ---- the other kernels of the path, same inputs (explain `value`; not part of it) -------------
    def timed(fn, n=10):
        fn(); torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b_.record(); torch.cuda.synchronize()
        return a.elapsed_time(b_) / n
    def other_kernels():
        infer_ms = timed(lambda: R.wkv7_forward_infer_(*ins, y))                 # chunked tcgen05 forward (no_grad)
        DB, DSTEPS = 32, 64                                                       # config c4: 32 prompts, decode steps
        dstate = torch.zeros(LAYERS, DB, H, C, C, dtype=torch.float32, device=dev)
        dins = [d[n][:4, :DB // 4 * 1].reshape(DB, 1, H * C).contiguous() for n in "qwkvab"]
        dy_ = torch.empty(DB, 1, H * C, dtype=torch.bfloat16, device=dev)
        def decode_steps():
            for _ in range(DSTEPS):
                for l in range(LAYERS):
                    R.wkv7_state_forward_(DB, 1, H * C, H, dstate[l], *dins, dy_)
        decode_ms = timed(decode_steps, n=2) / DSTEPS                            # WKV part of one decode step, 24 layers

        return infer_ms, decode_ms, DB

    def fused_leg():
        # ---- fused time-mix elementwise kernels at the same config ([8,4096,1024] activations; explain, not part of `value`)
        from rwkvtts_b200 import fused as FU
        CC = H * C
        act = lambda: torch.randn(B, T, CC, device=dev).bfloat16()
        par = lambda *sh: (0.5 * torch.randn(*sh, device=dev)).bfloat16()
        fx, fdo = act(), [act() for _ in range(6)]
        mixes = [par(1, 1, CC).requires_grad_(True) for _ in range(6)]
        k_, v_, wl_, al_, vl_, vf_ = (act().requires_grad_(True) for _ in range(6))
        pp = [par(1, 1, CC).requires_grad_(True) for _ in range(5)]
        y_, r_, g_ = (act().requires_grad_(True) for _ in range(3))
        rk_, lw_, lb_ = par(H, C).requires_grad_(True), par(CC).requires_grad_(True), par(CC).requires_grad_(True)
        fxg = fx.clone().requires_grad_(True)

        def fused_times():
            res = {}
            def fb(name, fwd, inputs, douts, n_fwd_arrays, n_bwd_arrays):
                outs = fwd()
                outs = outs if isinstance(outs, (tuple, list)) else (outs,)
                tf = timed(fwd, n=5)
                # autograd.grad: no accumulation into .grad, so the timed region is the backward kernels + a few allocations
                tb = timed(lambda: torch.autograd.grad(outs, inputs, douts[:len(outs)], retain_graph=True), n=5)
                nbytes = B * T * CC * 2
                res[name] = {"fwd_ms": tf, "bwd_ms": tb, "fwd_GBps": n_fwd_arrays * nbytes / tf / 1e6,
                             "bwd_GBps": n_bwd_arrays * nbytes / tb / 1e6,
                             "fwd_frac": n_fwd_arrays * nbytes / tf / 1e6 / peak_, "bwd_frac": n_bwd_arrays * nbytes / tb / 1e6 / peak_}
            peak_ = peaks()[0]
            fb("shift_mix6", lambda: FU.shift_mix(fxg, mixes), [fxg] + mixes, fdo, 7, 8)     # fwd 1r+6w; bwd 6r+1r(x)+1w
            fb("prep", lambda: FU.prep(k_, v_, wl_, al_, vl_, vf_, *pp), [k_, v_, wl_, al_, vl_, vf_] + pp, fdo, 11, 17)  # 6r+5w; 11r+6w
            fb("out", lambda: FU.out(y_, r_, k_, v_, g_, rk_, lw_, lb_, 64e-5), [y_, r_, k_, v_, g_, rk_, lw_, lb_], fdo, 6, 11)  # 5r+1w; 6r+5w
            res["note"] = ("algorithmic arrays of [B,T,C] bf16 moved per call / CUDA-event time of the autograd call (includes "
                           "the partial-sum reduce kernel and output allocations)")
            return res
        return fused_times()


    infer_ms = decode_ms = fused_k = None
    DB = 32
    if extras:
        log("leg: other kernels")
        infer_ms, decode_ms, DB = other_kernels()
        log("leg: fused kernels")
        fused_k = fused_leg()
        torch.cuda.empty_cache()
    log("leg: e2e")

    # ---- e2e: public API (WindBackstepping autograd op) with pinned HOST buffers --------------
    host_in = [x[n].pin_memory() for n in "wqkvab"] + [x["dy"].pin_memory()]
    host_out = [torch.empty_like(x["v"]).pin_memory() for _ in range(7)]
    h2d = sum(t_.numel() * t_.element_size() for t_ in host_in) * LAYERS
    d2h = sum(t_.numel() * t_.element_size() for t_ in host_out) * LAYERS

    # Copies of layer l+1 (H2D) and of layer l-1 (D2H) overlap the kernels of layer l: three streams, device input
    # buffers double buffered, events for the hand-offs.  PCIe is full duplex, so the step is bound by the larger of
    # the two copy directions (0.47 GB each way per layer), not by their sum.
    s_comp = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    dev_in = [[torch.empty_like(d["v"]) for _ in range(7)] for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]          # inputs of buffer b have landed
    ev_free = [torch.cuda.Event() for _ in range(2)]        # kernels are done with buffer b
    ev_out = [torch.cuda.Event() for _ in range(2)]         # results of buffer b are in host memory
    used = [False, False]

    def e2e_step():
        for l in range(LAYERS):
            b = l % 2
            with torch.cuda.stream(s_in):
                if l >= 2 or used[b]:
                    s_in.wait_event(ev_free[b])     # the kernels of the layer that last used this buffer are done
                for t_, h_ in zip(dev_in[b], host_in):
                    t_.copy_(h_, non_blocking=True)
                ev_in[b].record(s_in)
            s_comp.wait_event(ev_in[b])
            leaves = [t_.detach().requires_grad_(True) for t_ in dev_in[b][:6]]
            yy = R.WindBackstepping.apply(*leaves)
            yy.backward(dev_in[b][6])
            ev_free[b].record(s_comp)
            used[b] = True
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_free[b])
                outs = [yy.detach()] + [l_.grad for l_ in leaves]
                for o, r_ in zip(host_out, outs):
                    r_.record_stream(s_out)
                    o.copy_(r_, non_blocking=True)
                ev_out[b].record(s_out)
        s_comp.wait_stream(s_out)

    e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = dist_max(e0.elapsed_time(e1), world)
    e2e_value = B * T * world / (e2e_ms / args.e2e_steps * 1e-3)

    log("e2e done")
    if rank != 0:
        dist_finish(world)
        return

    # ---- reference CUDA op on the same inputs (extra data point; oracle/_ref) --------------------
    ref_gpu = None
    if extras and not args.no_ref_gpu:
        try:
            from oracle import c_oracle as CO
            if CO.ref_available():
                for _ in range(2):
                    yr, sr, sar = CO.ref_forward(*ins)
                    CO.ref_backward(*ins, d["dy"], sr, sar)
                torch.cuda.synchronize()
                r0, r1, r2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                n = 3
                r0.record()
                for _ in range(n):
                    yr, sr, sar = CO.ref_forward(*ins)
                r1.record()
                for _ in range(n):
                    CO.ref_backward(*ins, d["dy"], sr, sar)
                r2.record()
                torch.cuda.synchronize()
                rf, rb = r0.elapsed_time(r1) / n, r1.elapsed_time(r2) / n
                ref_gpu = {"what": "unmodified reference wind_backstepping kernels (oracle/_ref), same inputs, 1 GPU",
                           "fwd_ms": rf, "bwd_ms": rb, "tokens_per_s": B * T / ((rf + rb) * LAYERS * 1e-3),
                           "speedup_fwd": rf / fwd_ms, "speedup_bwd": rb / bwd_ms,
                           "speedup_step": (rf + rb) / (fwd_ms + bwd_ms)}
                del yr, sr, sar
        except Exception as e:                                  # never let the extra leg kill the line
            ref_gpu = {"error": repr(e)}

    peak, peak_src = peaks()
    th = B * T * H
    dom = "bwd" if bwd_ms >= fwd_ms else "fwd"
    dom_ms, dom_bytes = (bwd_ms, BWD_BYTES) if dom == "bwd" else (fwd_ms, FWD_BYTES)
    ach = dom_bytes * th / (dom_ms * 1e-3) / 1e9
    fwd_ach = FWD_BYTES * th / (fwd_ms * 1e-3) / 1e9
    bwd_ach = BWD_BYTES * th / (bwd_ms * 1e-3) / 1e9
    kernels = {"fwd_ms": fwd_ms, "bwd_ms": bwd_ms, "fwd_GBps": fwd_ach, "bwd_GBps": bwd_ach,
               "fwd_frac": fwd_ach / peak, "bwd_frac": bwd_ach / peak,
               "note": "fwd/bwd = training pair (chunked tcgen05 kernels, default family)"}
    if extras:
        kernels.update({
            "fwd_infer_tcgen05_ms": infer_ms, "fwd_infer_GBps": FWD_BYTES * th / (infer_ms * 1e-3) / 1e9,
            "fwd_infer_frac": FWD_BYTES * th / (infer_ms * 1e-3) / 1e9 / peak,
            "decode_step_wkv_ms": decode_ms,
            "decode_step_GBps": (2 * C * C * 4 + FWD_BYTES) * DB * H * LAYERS / (decode_ms * 1e-3) / 1e9,
            "decode_wkv_tokens_per_s": DB / (decode_ms * 1e-3),
            "note": "fwd/bwd = training pair (chunked tcgen05 kernels, default family); fwd_infer = snapshot-free tcgen05 "
                    "forward used under no_grad; decode = single-step kernel called eagerly, T=1, B=32, 24 layers (launch-bound)"})
    # DRAM bytes per launch of the two training kernels: dram__bytes_read.sum + dram__bytes_write.sum from the ncu
    # `--set full` capture of this same configuration (profiles/r01_tc_pair_full_v5.txt); the excess over the algorithmic
    # bytes is the checkpoint / U scratch the pair exchanges (537 + 134 MB written by the forward, read by the backward)
    NCU_TRAFFIC = {"fwd": 0.402706e9 + 0.679393e9, "bwd": 1.201027e9 + 0.380507e9}
    line = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16 in/out, fp32 state", "data": "synthetic",
        "config": {"workload": WORKLOAD, "B_per_gpu": B, "T": T, "H": H, "head": C, "layers": LAYERS,
                   "l2": "inputs+outputs of one layer-call are 0.47-0.87 GB, larger than the 126 MB L2; "
                         "no explicit flush"},
        "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "rwkvtts_b200.WindBackstepping (autograd) with pinned host tensors; H2D / kernels / D2H of consecutive layers overlapped on three streams", "steps": args.e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": f"wkv7 {dom}", "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "traffic": NCU_TRAFFIC[dom],
                     "traffic_source": "ncu --set full capture, profiles/r01_tc_pair_full_v5.txt", "peak_source": peak_src,
                     "algorithmic_bytes_per_token_head": dom_bytes, "token_heads_per_launch": th,
                     "avg_launch_ms": dom_ms},
        "kernels": kernels,
        "fused_tmix_kernels": fused_k,
        "ref_gpu_op": ref_gpu,
    }
    if world == 1 and not args.no_decode:
        log("leg: decode")
        try:
            line["decode"] = decode_leg(dev)
        except Exception as e:                                  # never let the extra leg kill the line
            line["decode"] = {"error": repr(e)}
    if world == 1 and not args.no_model_step:
        log("leg: model step")
        try:
            line["model_train_step"] = model_step_leg(dev)
        except Exception as e:
            line["model_train_step"] = {"error": repr(e)}
    if world == 1 and not args.no_cpu_baseline:
        log("leg: cpu baseline")
        line["cpu_baseline"] = cpu_reference()
    log("done")
    print(json.dumps(line))
    dist_finish(world)


if __name__ == "__main__":
    main()
