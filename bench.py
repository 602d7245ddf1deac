#!/usr/bin/env python
"""bench.py -- RWKV-7 0.4B Spark-layout training step on B200 (BASELINE.json configs[1] at N = 1, configs[2] at N = 8).

A "step" is one optimizer step of the reference's training loop (train_scripts/train_spark_rwkv7speech.py:664-691)
on one synthetic batch per rank (batch 8 x 4096 tokens):

    process_single_batch (embedding gather / concat of the Spark layout)  ->  engine(inputs_embeds, attention_mask,
    labels) (24 blocks: fused time-mix kernels, WKV-7 tcgen05 forward, cuBLAS GEMMs; fused linear + CE head)  ->
    NaN all-reduce  ->  engine.backward(loss)  ->  engine.step() (ZeRO-2: per-bucket NCCL reduce-scatter(AVG)
    overlapped with backward, fused Adam on the rank's slices, in-place all-gather)

through the DeepSpeed-compatible surface (`deepspeed.initialize` of this repo's shim).  metric = audio-tokens/s =
B*T*n_gpus / step time, ranks are data parallel with their own batch ("scaling": "weak").

  python bench.py [--gpus N --steps K --warmup W]    our arm: `value` with the token ids resident on the device,
                                                     `e2e` with ids coming from pinned host memory and the loss read
                                                     back every step; `roofline` for the dominant kernel of the hot
                                                     path (WKV-7 backward), timed inside the same region with CUDA events
  python bench.py --impl reference ...               the reference algorithm of the hot path on the host cores
                                                     (C port of wkv7_cuda.cu, OpenMP): a bounded, EXTRAPOLATED sample
  python bench.py --leg NAME                         one explanatory leg on its own (what the main run spawns, each in
                                                     its own process under its own timeout, at N = 1)

Under torchrun (N > 1) every rank runs; rank 0 prints the single JSON line.  Progress goes to stderr, one line per leg.
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
WORKLOAD = ("configs[1]: RWKV-7 0.4B Spark-layout bf16, batch 8/GPU, seq_len 4096, whole train step "
            "(batch builder, forward, backward, ZeRO-2 optimizer step)")
_T0 = time.time()


def log(msg):
    """One stderr line per leg: a hang can no longer erase the record of how far the run got."""
    sys.stderr.write("bench[%s +%.1fs]: %s\n" % (os.environ.get("RANK", "0"), time.time() - _T0, msg))
    sys.stderr.flush()


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


# =====================================================================================================================
# the train step (configs[1] / configs[2])
# =====================================================================================================================
def random_init_0p4b():
    """RWKV7ForSpeech 0.4B with the reference init, except that the projections the reference zero-initialises
    (rwkv_s2s_single_ffn.py:113,:156,:221) get small random values: zeros there would hide errors of the recurrence."""
    import torch
    from rwkvtts_b200.spark import RWKV7ForSpeech, spark_0p4b_config
    torch.manual_seed(42)
    model = RWKV7ForSpeech(spark_0p4b_config())
    with torch.no_grad():
        for _, p in model.named_parameters():
            if p.abs().sum() == 0:
                p.copy_(torch.randn_like(p) * 0.02)
    return model.to(torch.bfloat16)


def build_engine(world, fused_p2p=False):
    """Model, optimizer and engine as the reference's script builds them: configure_optimizer
    (train_spark_rwkv7speech.py:178-197: one group, FusedAdam betas (0.9, 0.95), eps 1e-18, adam_w_mode),
    ds_config with bf16 + ZeRO stage 2 + reduce_scatter (:483-516), deepspeed.initialize (:566-572)."""
    import deepspeed
    from deepspeed.ops.adam import FusedAdam
    model = random_init_0p4b()
    model.train()
    groups = [{"params": [p for _, p in sorted(model.named_parameters()) if p.requires_grad], "weight_decay": 0.0,
               "my_lr_scale": 1.0}]
    opt = FusedAdam(groups, lr=1e-4, betas=(0.9, 0.95), eps=1e-18, bias_correction=True, adam_w_mode=True,
                    amsgrad=False, weight_decay=0.01)
    cfg = {"distributed_backend": "nccl", "train_batch_size": B * world, "bf16": {"enabled": True},
           "zero_optimization": {"stage": 2, "allgather_partitions": True, "reduce_scatter": True,
                                 "overlap_comm": True, "contiguous_gradients": True, "fused_p2p": bool(fused_p2p)},
           "gradient_checkpointing": False, "dump_state": False}
    engine, _, _, _ = deepspeed.initialize(model=model, config=cfg, model_parameters=model.parameters(), optimizer=opt)
    return engine


def train_arm(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "INFO"                  # leave NCCL's log on: communicator size is checkable
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the one JSON line
        import deepspeed
        log("joining the NCCL process group")
        deepspeed.init_distributed("nccl")
        dist.barrier(device_ids=[local_rank])
        log("process group up (%d ranks)" % dist.get_world_size())
    import rwkvtts_b200 as R
    from rwkvtts_b200 import ops
    from rwkvtts_b200.batch import process_single_batch
    from rwkvtts_b200.spark import synthetic_spark_batch
    lib = R._lib.lib()
    log("building the 0.4B model and the engine")
    engine = build_engine(world, fused_p2p=args.zero_p2p)
    n_params = engine.numel
    host_batch = synthetic_spark_batch(B, T, seed=42 + rank)
    dev_batch = {k: v.to(dev) for k, v in host_batch.items()}
    nan_flag = torch.zeros(1, device=dev)

    def step(batch, read_loss=False):
        pb = process_single_batch(batch, engine, eos_token_id=8192)
        loss = engine(inputs_embeds=pb["input_embs"], attention_mask=pb["attention_mask"], labels=pb["labels"]).loss
        # the script's NaN guard (train_spark_rwkv7speech.py:664-670) without its host sync: the decision stays on the
        # device (a non-finite loss gives non-finite gradients, which engine.step() skips on every rank)
        if world > 1:
            nan_flag.copy_((~torch.isfinite(loss.detach())).float().reshape(1))
            dist.all_reduce(nan_flag, op=dist.ReduceOp.MAX)
        engine.backward(loss)
        engine.step()
        return float(loss) if read_loss else loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(device_ids=[local_rank])
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world <= 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W = max(args.warmup, 3)
    log("warm-up (%d steps)" % W)
    for i in range(W):
        l0 = step(dev_batch, read_loss=True)
        log("  warm-up step %d loss %.4f" % (i, l0))
    barrier()
    # ---- timed region: K steps, ids resident on the device --------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.TIMING = {"fwd": [], "bwd": []}
    launches0 = lib.rwkvtts_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    log("timed region (%d steps)" % args.steps)
    prof = os.environ.get("RWKVTTS_BENCH_PROFILE") == "1"        # ncu --profile-from-start off: only the timed region
    if prof:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(args.steps):
        loss = step(dev_batch)
    e1.record()
    barrier()
    if prof:
        torch.cuda.profiler.stop()
    launches = lib.rwkvtts_kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    timing, ops.TIMING = ops.TIMING, None
    ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    value = B * T * world / (ms_per_step * 1e-3)
    fwd_ms = sum(a.elapsed_time(b) for a, b in timing["fwd"]) / max(len(timing["fwd"]), 1)
    bwd_ms = sum(a.elapsed_time(b) for a, b in timing["bwd"]) / max(len(timing["bwd"]), 1)
    last_loss = float(loss)
    log("value %.0f tokens/s (%.2f ms/step), wkv fwd %.3f ms bwd %.3f ms per launch" % (value, ms_per_step, fwd_ms, bwd_ms))
    # ---- e2e: ids from pinned host memory every step, loss read back every step -----------------------------------
    h2d = sum(v.numel() * v.element_size() for v in host_batch.values())

    def e2e_step():
        b = {k: v.to(dev, non_blocking=True) for k, v in host_batch.items()}
        return step(b, read_loss=True)          # float(loss): 4-byte D2H + host sync, as the script's loss.item()

    e2e_step()
    barrier()
    n_e2e = max(2, min(args.steps, args.e2e_steps))
    e0.record()
    for _ in range(n_e2e):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / n_e2e
    e2e_value = B * T * world / (e2e_ms * 1e-3)
    log("e2e %.0f tokens/s" % e2e_value)
    # ---- the exchange in isolation (collective's share of the step) -----------------------------------------------
    comm = engine.profile_comm() if world > 1 else None
    if comm is not None:
        comm["share_of_step_if_not_overlapped"] = ((comm.get("fused_exchange_and_adam_ms") or 0.0) + comm["reduce_scatter_ms"]
                                                   + comm["all_gather_ms"]) / ms_per_step
        comm["nccl_nranks"] = dist.get_world_size()
        comm["what"] = comm.get("mode") or (
            "per step and rank: reduce-scatter(AVG) of %d buckets of bf16 gradients, one 2-float all-reduce, all-gather of "
            "the updated bf16 parameters; timed back to back without the step around them" % comm["buckets"])
    if world > 1:
        dist.barrier(device_ids=[local_rank])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    th = B * T * H
    dom = "bwd" if bwd_ms >= fwd_ms else "fwd"
    dom_ms, dom_bytes = (bwd_ms, BWD_BYTES) if dom == "bwd" else (fwd_ms, FWD_BYTES)
    ach = dom_bytes * th / (dom_ms * 1e-3) / 1e9
    traffic = load_ncu_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 (fp32 recurrent state, fp32 master weights)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "B_per_gpu": B, "T": T, "H": H, "head": C, "layers": LAYERS,
                   "params": n_params, "parallelism": "dp%d (ZeRO-2: reduce-scatter / sharded Adam / all-gather)" % world,
                   "l2": "activations of one layer are 67 MB per tensor, ~14 live per layer, far beyond the 126 MB L2; "
                         "no explicit flush"},
        "loss": last_loss,
        "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 4 * world,
                "api": "deepspeed.initialize -> process_single_batch(host ids, pinned) -> engine(**batch) -> "
                       "engine.backward -> engine.step -> float(loss)", "steps": n_e2e},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "wkv7_tc_%s_kernel" % dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "traffic": traffic.get(dom), "traffic_source": traffic.get("source"),
                     "peak_source": peak_src, "algorithmic_bytes_per_token_head": dom_bytes, "token_heads_per_launch": th,
                     "avg_launch_ms": dom_ms, "launches_timed": len(timing[dom]),
                     "how": "CUDA events around every launch of the kernel inside the timed region of the train step"},
        "kernels": {"wkv_fwd_ms": fwd_ms, "wkv_bwd_ms": bwd_ms,
                    "wkv_fwd_GBps": FWD_BYTES * th / (fwd_ms * 1e-3) / 1e9, "wkv_bwd_GBps": BWD_BYTES * th / (bwd_ms * 1e-3) / 1e9,
                    "wkv_fwd_frac": FWD_BYTES * th / (fwd_ms * 1e-3) / 1e9 / peak,
                    "wkv_bwd_frac": BWD_BYTES * th / (bwd_ms * 1e-3) / 1e9 / peak,
                    "wkv_share_of_step": (fwd_ms + bwd_ms) * LAYERS / ms_per_step},
        "collective": comm,
    }
    if world > 1:
        dist.destroy_process_group()
    # ---- explanatory legs, each in its own process under its own timeout (N = 1 only) -----------------------------
    if world == 1 and not args.no_legs:
        del engine
        torch.cuda.empty_cache()
        for name, tmo in LEGS:
            if name in args.skip:
                continue
            line[name] = run_leg(name, tmo)
    print(json.dumps(line))
    log("done")


def load_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the two training kernels, from the round's committed
    ncu --set full capture (profiles/r02_wkv_traffic.json, written by scripts/ncu_summary.py from the capture of this
    configuration); null when the round has no capture for the kernels as built."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_wkv_traffic.json")))
        return {"fwd": t.get("fwd_bytes"), "bwd": t.get("bwd_bytes"), "source": t.get("source")}
    except Exception:
        return {"fwd": None, "bwd": None, "source": None}


def run_leg(name, timeout_s):
    log("leg: %s (own process, timeout %d s)" % (name, timeout_s))
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "OMP_NUM_THREADS"):     # torchrun's exports are not for the legs
        env.pop(k, None)
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--leg", name], capture_output=True, text=True,
                           timeout=timeout_s, env=env)
    except subprocess.TimeoutExpired:
        return {"error": "leg exceeded its %d s limit and was killed" % timeout_s}
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode != 0 or not lines:
        return {"error": "leg exited %d" % r.returncode, "stderr_tail": r.stderr[-600:]}
    try:
        return json.loads(lines[-1])
    except ValueError:
        return {"error": "leg printed no JSON", "stdout_tail": r.stdout[-300:]}


# =====================================================================================================================
# explanatory legs (python bench.py --leg NAME)
# =====================================================================================================================
def _timed(fn, n=10):
    import torch
    fn(); torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b_.record(); torch.cuda.synchronize()
    return a.elapsed_time(b_) / n


def _op_inputs(dev, b=B, t=T, h=H, seed=42):
    import torch
    from rwkvtts_b200.synth import make_inputs
    x = make_inputs(b, t, h, seed=seed)
    d = {n: v.to(dev) for n, v in x.items()}
    ins = [d[n] for n in "wqkvab"]
    y = torch.empty_like(d["v"])
    s = torch.empty(b, h, t // 16, C, C, dtype=torch.float32, device=dev)
    sa = torch.empty(b, t, h, C, dtype=torch.float32, device=dev)
    grads = [torch.empty_like(d["v"]) for _ in range(6)]
    return d, ins, y, s, sa, grads


def leg_wkv_ops():
    """The recurrence alone on resident [8,4096,16,64] tensors (round 1's headline): forward + backward per layer-call,
    the snapshot-free forward and the decode-step kernel; and the same pair at config c5's per-GPU shape
    [2,8192,32,64] over T = 1k..8k."""
    import torch
    import rwkvtts_b200 as R
    dev = torch.device("cuda", 0)
    peak, _ = peaks()
    d, ins, y, s, sa, grads = _op_inputs(dev)
    th = B * T * H
    f = _timed(lambda: R.wkv7_forward_(*ins, y, s, sa))
    bw = _timed(lambda: R.wkv7_backward_(*ins, d["dy"], s, sa, *grads))
    inf = _timed(lambda: R.wkv7_forward_infer_(*ins, y))
    out = {"shape": [B, T, H, C], "fwd_ms": f, "bwd_ms": bw, "fwd_infer_ms": inf,
           "fwd_frac": FWD_BYTES * th / (f * 1e-3) / 1e9 / peak, "bwd_frac": BWD_BYTES * th / (bw * 1e-3) / 1e9 / peak,
           "fwd_infer_frac": FWD_BYTES * th / (inf * 1e-3) / 1e9 / peak,
           "tokens_per_s_24_layers_pair_only": B * T / ((f + bw) * LAYERS * 1e-3)}
    DB = 32
    dstate = torch.zeros(LAYERS, DB, H, C, C, dtype=torch.float32, device=dev)
    dins = [d[n][:4, :DB // 4].reshape(DB, 1, H * C).contiguous() for n in "qwkvab"]
    dy_ = torch.empty(DB, 1, H * C, dtype=torch.bfloat16, device=dev)

    def decode_steps():
        for l in range(LAYERS):
            R.wkv7_state_forward_(DB, 1, H * C, H, dstate[l], *dins, dy_)
    dms = _timed(decode_steps, n=20)
    out["decode_step_wkv_ms_24_layers_eager"] = dms
    out["decode_step_GBps"] = (2 * C * C * 4 + FWD_BYTES) * DB * H * LAYERS / (dms * 1e-3) / 1e9
    del d, ins, y, s, sa, grads
    torch.cuda.empty_cache()
    c5 = []
    for t5 in (1024, 2048, 4096, 8192):
        d, ins, y, s, sa, grads = _op_inputs(dev, 2, t5, 32, seed=5)
        th5 = 2 * t5 * 32
        f5 = _timed(lambda: R.wkv7_forward_(*ins, y, s, sa))
        b5 = _timed(lambda: R.wkv7_backward_(*ins, d["dy"], s, sa, *grads))
        c5.append({"T": t5, "fwd_ms": f5, "bwd_ms": b5, "fwd_GBps": FWD_BYTES * th5 / (f5 * 1e-3) / 1e9,
                   "bwd_GBps": BWD_BYTES * th5 / (b5 * 1e-3) / 1e9, "fwd_frac": FWD_BYTES * th5 / (f5 * 1e-3) / 1e9 / peak,
                   "bwd_frac": BWD_BYTES * th5 / (b5 * 1e-3) / 1e9 / peak})
        del d, ins, y, s, sa, grads
    out["c5_per_gpu_shape_[2,T,32,64]"] = c5
    return out


def leg_ref_gpu():
    """The unmodified reference CUDA kernels (oracle/_ref, compiled from /root/reference/model/llm/cuda) on the same
    inputs: op level, and R-GPU-model -- the reference's eager time-mix / channel-mix chain
    (rwkv_s2s_single_ffn.py:158-196, :223-230 as restated in core.tmix with FUSED off) around the reference op."""
    import torch
    import rwkvtts_b200 as R
    from oracle import c_oracle as CO
    if not CO.ref_available():
        return {"error": "oracle/_ref not built"}
    dev = torch.device("cuda", 0)
    d, ins, y, s, sa, grads = _op_inputs(dev)
    f = _timed(lambda: R.wkv7_forward_(*ins, y, s, sa), n=5)
    bw = _timed(lambda: R.wkv7_backward_(*ins, d["dy"], s, sa, *grads), n=5)
    st = {}

    def rf():
        st["y"], st["s"], st["sa"] = CO.ref_forward(*ins)
    rf_ms = _timed(rf, n=3)
    rb_ms = _timed(lambda: CO.ref_backward(*ins, d["dy"], st["s"], st["sa"]), n=3)
    out = {"what": "unmodified reference wind_backstepping kernels (oracle/_ref), same inputs, 1 GPU",
           "fwd_ms": rf_ms, "bwd_ms": rb_ms, "ours_fwd_ms": f, "ours_bwd_ms": bw,
           "tokens_per_s_24_layers_pair_only": B * T / ((rf_ms + rb_ms) * LAYERS * 1e-3),
           "speedup_fwd": rf_ms / f, "speedup_bwd": rb_ms / bw, "speedup_pair": (rf_ms + rb_ms) / (f + bw)}
    del d, ins, y, s, sa, grads, st
    torch.cuda.empty_cache()
    try:
        out["model"] = _ref_gpu_model(dev)
    except Exception as e:                                  # the op-level numbers stand on their own
        out["model"] = {"error": repr(e)}
    return out


def _ref_gpu_model(dev, steps=2):
    """R-GPU-model: whole 0.4B model forward + backward with the ATen elementwise chain of the
    reference's Block and the reference's own CUDA op inside, against the same model on this repo's kernels."""
    import torch
    from oracle import c_oracle as CO
    from rwkvtts_b200 import core

    class RefOp(torch.autograd.Function):       # WindBackstepping with the reference kernels (rwkv_s2s_single_ffn.py:15-35)
        @staticmethod
        def forward(ctx, w, q, k, v, z, b):
            y, s, sa = CO.ref_forward(w, q, k, v, z, b)
            ctx.save_for_backward(w, q, k, v, z, b, s, sa)
            return y

        @staticmethod
        def backward(ctx, dy):
            w, q, k, v, z, b, s, sa = ctx.saved_tensors
            return tuple(CO.ref_backward(w, q, k, v, z, b, dy.contiguous(), s, sa))

    m = random_init_0p4b().to(dev).train()
    ids = torch.randint(0, 8192, (B, T), device=dev)
    res = {}
    try:
        for name, fused_on, op in (("ours_fused", True, None), ("reference_op_aten_chain", False, RefOp)):
            core.FUSED = fused_on
            core.WKV_TRAIN_OP = op

            def one():
                m.zero_grad(set_to_none=True)
                m(input_ids=ids, labels=ids).loss.backward()
            ms = _timed(one, n=steps)
            res[name] = {"ms_fwd_bwd": ms, "tokens_per_s": B * T / ms * 1e3}
    finally:
        core.FUSED = True
        core.WKV_TRAIN_OP = None
    res["speedup"] = res["reference_op_aten_chain"]["ms_fwd_bwd"] / res["ours_fused"]["ms_fwd_bwd"]
    res["what"] = ("0.4B model, batch 8 x 4096, forward + backward, no optimizer: reference CUDA op + ATen elementwise chain "
                   "(the reference's eager Block) vs this repo's kernels; GEMMs are cuBLAS in both")
    return res


def leg_fused():
    """Fused time-mix elementwise kernels at [8,4096,1024] through autograd (includes the reduce kernel)."""
    import torch
    from rwkvtts_b200 import fused as FU
    dev = torch.device("cuda", 0)
    CC = H * C
    peak_ = peaks()[0]
    act = lambda: torch.randn(B, T, CC, device=dev).bfloat16()
    par = lambda *sh: (0.5 * torch.randn(*sh, device=dev)).bfloat16()
    fx, fdo = act(), [act() for _ in range(6)]
    mixes = [par(1, 1, CC).requires_grad_(True) for _ in range(6)]
    k_, v_, wl_, al_, vl_, vf_ = (act().requires_grad_(True) for _ in range(6))
    pp = [par(1, 1, CC).requires_grad_(True) for _ in range(5)]
    y_, r_, g_ = (act().requires_grad_(True) for _ in range(3))
    rk_, lw_, lb_ = par(H, C).requires_grad_(True), par(CC).requires_grad_(True), par(CC).requires_grad_(True)
    fxg = fx.clone().requires_grad_(True)
    res = {}

    def graph_time(fn, n=10):
        """fn captured in a CUDA graph and replayed: the GPU-side time of the call's kernels (parameter conversions, scratch
        fills and the partial-sum reduction included) without the launch gaps of an isolated eager loop -- in the train step
        the queue is deep and those gaps do not exist."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return _timed(g.replay, n=n)

    def fb(name, fwd, inputs, douts, n_fwd_arrays, n_bwd_arrays):
        def as_tuple(o):
            return o if isinstance(o, (tuple, list)) else (o,)
        def eager():
            outs = as_tuple(fwd())
            return (len(outs), _timed(fwd, n=5),
                    _timed(lambda: torch.autograd.grad(outs, inputs, douts[:len(outs)], retain_graph=True), n=5))
        n_out, tf_e, tb_e = eager()        # (its autograd graph dies here: a live one pins the leaves' AccumulateGrad nodes to
        import gc                          # the default stream, which invalidates a capture on another stream)
        gc.collect()
        outs = [None] * n_out
        nbytes = B * T * CC * 2
        res[name] = {"fwd_ms_eager_loop": tf_e, "bwd_ms_eager_loop": tb_e}
        try:
            tf = graph_time(fwd)
            tfb = graph_time(lambda: torch.autograd.grad(as_tuple(fwd()), inputs, douts[:len(outs)]))
            tb = max(tfb - tf, 1e-6)
            res[name]["timing"] = "CUDA-graph replay (backward = forward+backward - forward)"
        except Exception as e:                                    # capture refused: keep the eager-loop figures
            tf, tb = tf_e, tb_e
            res[name]["timing"] = "eager loop (graph capture failed: %s)" % str(e)[:120]
        res[name].update({"fwd_ms": tf, "bwd_ms": tb, "fwd_frac": n_fwd_arrays * nbytes / tf / 1e6 / peak_,
                          "bwd_frac": n_bwd_arrays * nbytes / tb / 1e6 / peak_})
    fb("shift_mix6", lambda: FU.shift_mix(fxg, mixes), [fxg] + mixes, fdo, 7, 8)
    fb("prep", lambda: FU.prep(k_, v_, wl_, al_, vl_, vf_, *pp), [k_, v_, wl_, al_, vl_, vf_] + pp, fdo, 11, 17)
    fb("out", lambda: FU.out(y_, r_, k_, v_, g_, rk_, lw_, lb_, 64e-5), [y_, r_, k_, v_, g_, rk_, lw_, lb_], fdo, 6, 11)
    lx_, lr_ = act().requires_grad_(True), act().requires_grad_(True)
    fb("add_layernorm", lambda: FU.add_layernorm(lx_, lr_, lw_, lb_, 1e-5), [lx_, lr_, lw_, lb_], fdo, 4, 4)
    res["note"] = ("algorithmic [B,T,C] bf16 arrays moved per call / CUDA-event time of the autograd call (all of its kernels), "
                   "frac of measured HBM peak; *_eager_loop = the same call timed in an isolated eager loop, where launch gaps count")
    return res


def leg_decode():
    """AR decode, BASELINE config c4: 0.4B, 32 prompts of 163 positions, greedy, EOS suppressed, 2000 new tokens through
    RWKV7ForCausalLM.generate (CUDA-graph step); and the north_star's parity criterion: the greedy ids against a decode
    loop built on the reference's own step kernel (oracle/_ref libref_state_fwd, rwkv7_state_fwd_fp16.cu) around the
    same ATen chain (scripts/decode_parity.py)."""
    import torch
    dev = torch.device("cuda", 0)
    DB, PROMPT, NEW = 32, 163, 2000
    m = random_init_0p4b().to(dev).eval()
    torch.manual_seed(7)
    ids = torch.randint(0, 8192, (DB, PROMPT), device=dev)
    kw = dict(input_ids=ids, do_sample=False, eos_token_id=None, use_cuda_graph=True)
    m.generate(max_new_tokens=10, **kw)
    torch.cuda.synchronize()
    short = 16                     # both runs pay prefill (+ graph capture): the difference is pure decode steps
    t0 = time.perf_counter(); m.generate(max_new_tokens=short, **kw); torch.cuda.synchronize()
    t_short = time.perf_counter() - t0
    t0 = time.perf_counter(); seq = m.generate(max_new_tokens=NEW, **kw); torch.cuda.synchronize()
    dt = time.perf_counter() - t0 - t_short
    from rwkvtts_b200.decode import unsupported_reason
    out = {"tokens_per_s": DB * (NEW - short) / dt, "ms_per_step": dt / (NEW - short) * 1e3, "steps": NEW,
           "step": "one persistent kernel per token over all layers + head + device arg-max (csrc/decode_step.cu)"
                   if unsupported_reason(m, DB) is None else "CUDA-graph step",
           "prefill_plus_setup_ms": (t_short - short * dt / (NEW - short)) * 1e3,
           "config": "configs[3]: RWKV-7 0.4B random init, batch 32, prompt 163, greedy, EOS suppressed, 2000 new tokens; "
                     "steps timed as the difference of the 2000-token and a 16-token run"}
    # the step it replaces: the same kernels per layer as ~430 nodes of one CUDA graph
    kwg = dict(kw, use_megakernel=False)
    m.generate(max_new_tokens=short, **kwg); torch.cuda.synchronize()
    t0 = time.perf_counter(); m.generate(max_new_tokens=short, **kwg); torch.cuda.synchronize()
    tg_short = time.perf_counter() - t0
    t0 = time.perf_counter(); seqg = m.generate(max_new_tokens=NEW, **kwg); torch.cuda.synchronize()
    dtg = time.perf_counter() - t0 - tg_short
    out["graph_step"] = {"tokens_per_s": DB * (NEW - short) / dtg, "ms_per_step": dtg / (NEW - short) * 1e3}
    # HBM floor of one step: every weight once + every recurrent state in and out
    wbytes = sum(p.numel() * p.element_size() for n, p in m.named_parameters() if "embed" not in n)
    sbytes = 2 * DB * 24 * 16 * 64 * 64 * 4
    out["hbm_floor_ms"] = (wbytes + sbytes) / (peaks()[0] * 1e9) * 1e3
    out["frac_of_hbm_floor"] = out["hbm_floor_ms"] / out["ms_per_step"]
    try:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import decode_parity
        # exact mode (the reference's operations one for one, csrc/wkv7_step_exact.cu): must be identical
        kwx = dict(kw, exact=True)
        m.generate(max_new_tokens=short, **kwx); torch.cuda.synchronize()
        t0 = time.perf_counter(); m.generate(max_new_tokens=short, **kwx); torch.cuda.synchronize()
        tx_short = time.perf_counter() - t0
        t0 = time.perf_counter(); seqx = m.generate(max_new_tokens=NEW, **kwx); torch.cuda.synchronize()
        dtx = time.perf_counter() - t0 - tx_short
        out["exact_mode"] = {"tokens_per_s": DB * (NEW - short) / dtx, "ms_per_step": dtx / (NEW - short) * 1e3}
        out["greedy_ids_identical_vs_reference"] = decode_parity.compare(m, ids, seqx[:, PROMPT:], NEW)
        out["greedy_ids_identical_vs_reference"]["path"] = "generate(exact=True)"
        # the default fast path (fused kernels, fp32 intermediates): agreement rate and how close the reference's own
        # top-2 logits are wherever it picks another id
        out["fast_path_vs_reference"] = decode_parity.compare(m, ids, seq[:, PROMPT:], NEW)
        out["graph_step_vs_reference"] = decode_parity.compare(m, ids, seqg[:, PROMPT:], NEW)
    except Exception as e:
        out["greedy_ids_identical_vs_reference"] = {"error": repr(e)}
    return out


def cpu_reference(seconds_target=12.0):
    """The reference algorithm of the hot path (C port of forward_kernel/backward_kernel) on the host cores, on a bounded
    sample: one layer-call of [b,4096,16,64]; tokens/s EXTRAPOLATED to the 24-layer step (x 24 layers x B/b); the GEMMs,
    elementwise chain, loss and optimizer of the step are not in it (they would only make the CPU figure smaller)."""
    from oracle import c_oracle as CO
    from oracle import wkv7_oracle as O
    CO.build(ref=False)
    cores = CO.num_threads()
    b = min(B, max(1, -(-2 * cores // H)))      # at least 2 (b,h) units per thread
    x = O.make_inputs(b, T, H, seed=42)
    a = [x[n] for n in "wqkvab"]
    t0 = time.perf_counter()
    y, s, sa = CO.c_forward(*a)
    CO.c_backward(*a, x["dy"], s, sa)
    dt = time.perf_counter() - t0                       # warm-up + calibration
    reps = max(1, min(8, int(seconds_target / max(dt, 1e-3))))
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        y, s, sa = CO.c_forward(*a)
        CO.c_backward(*a, x["dy"], s, sa)
        best = min(best, time.perf_counter() - t0)
    value = b * T / (best * LAYERS)
    try:
        model = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        model = "unknown"
    return {"value": value, "unit": "tokens/s", "cores": cores, "kind": "port", "extrapolated": True,
            "sample": f"EXTRAPOLATED x({LAYERS} layers * B/b): 1 of {LAYERS} layers, batch {b} of {B}, T={T}, H={H}: WKV-7 "
                      f"fwd+bwd best of {reps} = {best * 1e3:.0f} ms; hot path only (no GEMMs / loss / optimizer); C port "
                      f"of wkv7_cuda.cu forward_kernel/backward_kernel, OpenMP over (b,h); cpu: {model}"}


LEGS = [("wkv_ops", 240), ("ref_gpu_op", 420), ("fused_tmix_kernels", 180), ("decode", 420), ("cpu_baseline", 120)]
LEG_FN = {"wkv_ops": leg_wkv_ops, "ref_gpu_op": leg_ref_gpu, "fused_tmix_kernels": leg_fused, "decode": leg_decode,
          "cpu_baseline": cpu_reference}


def run_reference(args, rank):
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone and may use every host
    # core (the other ranks exit without work), so undo that before the OpenMP runtime of the C port starts
    if "LOCAL_RANK" in os.environ or os.environ.get("OMP_NUM_THREADS") == "1":
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity")
                                            else (os.cpu_count() or 1))
    cb = cpu_reference(seconds_target=float(os.environ.get("RWKVTTS_BENCH_CPU_SECONDS", "20")))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": B * T / cb["value"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 in / fp32 state",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "cpu_arm": "EXTRAPOLATED x(24 layers * B/b) from a bounded sample of the "
                                                      "hot path (WKV-7 fwd+bwd) on the host cores"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--leg", default=None, choices=sorted(LEG_FN))
    ap.add_argument("--no-legs", action="store_true")
    ap.add_argument("--zero-p2p", action="store_true",
                    help="N > 1: fuse reduce-scatter + Adam + all-gather into one kernel over NVLink symmetric memory "
                         "(NVSwitch multicast when available) instead of the overlapped NCCL exchange")
    ap.add_argument("--skip", default=set(), type=lambda s: set(x for x in s.split(",") if x))
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.leg is not None:
        try:
            print(json.dumps(LEG_FN[args.leg]()))
        except RuntimeError as e:
            from rwkvtts_b200 import _lib
            print(json.dumps({"error": str(e).splitlines()[0], "watchdog": _lib.watchdog_report()}))
            sys.stdout.flush()
            os._exit(0)
        return
    try:
        train_arm(args, rank, local_rank, world)
    except RuntimeError:
        try:
            from rwkvtts_b200 import _lib
            rep = _lib.watchdog_report()
            if rep:
                log(rep)
        except Exception:
            pass
        raise


if __name__ == "__main__":
    main()
