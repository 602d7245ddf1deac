"""DeepSpeed-compatible training engine for the data-parallel path (SURVEY.md section 8b "engine surface", 8e).

Implements exactly what the reference's training scripts use (counted over /root/reference/train_scripts/):

    deepspeed.init_distributed()
    deepspeed.initialize(model=, config=dict, model_parameters=, optimizer=, lr_scheduler=)
        -> (engine, optimizer, None, lr_scheduler)                     train_spark_rwkv7speech_jsonl.py:428-433
    engine(**batch) / engine.backward(loss) / engine.step()            :479-482, :559-560
    engine.save_checkpoint(dir, tag=None) / load_checkpoint            :221, :598
    engine.train() / eval() / local_rank / device / module / parameters() / zero_grad() / generate(...)
    and attribute pass-through to the wrapped module (spark_dataset.py:175-203 passes the engine where a
    model is expected).
    deepspeed.ops.adam.FusedAdam / DeepSpeedCPUAdam(param_groups with 'my_lr_scale', 'name')   :195-199
    deepspeed.checkpointing.checkpoint(fn, *args)                      rwkv_s2s_single_ffn.py:315-316

Parallelism is what the reference gets from ZeRO stage 2 and nothing else: one process per GPU, the batch
sharded by the caller's DistributedSampler, no collective on the WKV / time-mix data path.  Once per
optimizer step:
    reduce-scatter (AVG) of the flat gradient buffer  ->  fused Adam on the rank's fp32 shard
    (rwkvtts_adam_shard, CUDA)  ->  all-gather of the updated parameters (in place, flat buffer),
plus one 2-float all-reduce for the gradient norm / non-finite check.  Parameters and gradients are views
into two flat buffers, so the collectives take no packing copies.  On one rank no collective runs.
"""
from __future__ import annotations

import ctypes
import math
import os
import warnings
from typing import Any, Dict, Iterable, List, Optional

import torch
import torch.distributed as dist

from . import _lib

ALIGN = 256     # elements; keeps every shard boundary 512-byte aligned for bf16


def _invalidate_param_cache():
    from . import fused              # late: fused pulls in the op registrations
    fused.invalidate_param_cache()


# ---------------------------------------------------------------------------------------------
# optimizers (deepspeed.ops.adam)
# ---------------------------------------------------------------------------------------------
class FusedAdam(torch.optim.Optimizer):
    """Holder of the param groups and hyper-parameters; the update itself is done by the engine on the
    rank's shard.  Used stand-alone (without an engine) it updates whole tensors with the same kernel."""

    def __init__(self, params, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-8, adam_w_mode=True,
                 weight_decay=0.0, amsgrad=False, adamw_mode=None, **kwargs):
        if amsgrad:
            raise RuntimeError("FusedAdam does not support the AMSGrad variant.")
        if adamw_mode is not None:           # DeepSpeedCPUAdam spells it adamw_mode
            adam_w_mode = adamw_mode
        defaults = dict(lr=lr, bias_correction=bias_correction, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.adam_w_mode = 1 if adam_w_mode else 0

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            group["step"] = group.get("step", 0) + 1
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["master"] = p.detach().float().clone()
                    st["exp_avg"] = torch.zeros_like(st["master"])
                    st["exp_avg_sq"] = torch.zeros_like(st["master"])
                adam_update(st["master"].view(-1), st["exp_avg"].view(-1), st["exp_avg_sq"].view(-1),
                            p.grad.contiguous().view(-1), p.data.view(-1), group, group["step"], self.adam_w_mode, 1.0)
        return loss


DeepSpeedCPUAdam = FusedAdam     # 180 GB of HBM per GPU: optimizer state stays on the device


def adam_update(master, m, v, grad, param, group: Dict[str, Any], step: int, adamw: int, gscale: float):
    """One Adam step on flat views.  CUDA tensors go through the library kernel; CPU tensors (the gloo
    tests of the host logic) use the same formula in torch."""
    b1, b2 = group["betas"]
    bc1 = 1.0 - b1 ** step if group.get("bias_correction", True) else 1.0
    bc2s = math.sqrt(1.0 - b2 ** step) if group.get("bias_correction", True) else 1.0
    lr, eps, wd = float(group["lr"]), float(group["eps"]), float(group["weight_decay"])
    n = master.numel()
    if n == 0:
        return
    _invalidate_param_cache()        # `param` is rewritten below without any parameter's version counter moving
    if master.is_cuda:
        for t, dt in ((grad, (torch.bfloat16, torch.float32)), (param, (torch.bfloat16, torch.float32))):
            if t.dtype not in dt:
                raise _lib.RwkvttsError(f"adam shard: unsupported dtype {t.dtype}")
        with torch.cuda.device(master.device):
            rc = _lib.lib().rwkvtts_adam_shard(
                master.data_ptr(), m.data_ptr(), v.data_ptr(), grad.data_ptr(), int(grad.dtype == torch.bfloat16),
                param.data_ptr(), int(param.dtype == torch.bfloat16), n, lr, b1, b2, eps, wd, adamw, bc1, bc2s,
                float(gscale), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "rwkvtts_adam_shard")
        return
    g = grad.float() * gscale
    if not adamw:
        g = g + wd * master
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    upd = (m / bc1) / (v.sqrt() / bc2s + eps)
    if adamw:
        upd = upd + wd * master
    master.add_(upd, alpha=-lr)
    param.copy_(master)


# ---------------------------------------------------------------------------------------------
# distributed helpers
# ---------------------------------------------------------------------------------------------
def init_distributed(dist_backend: Optional[str] = None, **kwargs) -> None:
    """deepspeed.init_distributed(): joins the torchrun rendezvous (RANK/WORLD_SIZE/MASTER_* env)."""
    if dist.is_initialized() or int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return
    backend = dist_backend or ("nccl" if torch.cuda.is_available() else "gloo")
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend=backend)


def _world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)


class Engine:
    """ZeRO-2 style data-parallel engine.  Flat layout (all in the model dtype):

        flat space  = optimizer parameters in REVERSE model order (the order their gradients become final in
                      backward), each padded to 8 elements, cut into buckets of ~`bucket_elems` at parameter
                      boundaries, each bucket padded to 256 * world_size elements;
        bucket b    = [start_b, end_b); rank r owns slice r of every bucket; the rank's fp32 master weights and Adam
                      moments are the concatenation of its slices ("shard space").

    Per optimizer step and bucket: reduce-scatter(AVG) of the bucket's gradients as soon as the last gradient of the
    bucket has been accumulated (post-accumulate-grad hooks; NCCL on a side stream, so the exchange of the last
    layers overlaps the backward of the first ones), then -- after one 2-float all-reduce of {squared norm,
    non-finite count} -- the fused Adam kernel on the slice (hyper-parameters per param group through a segment
    table, clipping / skip decided on the device) and the in-place all-gather of the updated parameters, the
    all-gather of bucket b overlapping the Adam kernel of bucket b+1.  No host synchronisation in step()."""

    def __init__(self, model: torch.nn.Module, optimizer: FusedAdam, config: Dict[str, Any],
                 lr_scheduler=None, device: Optional[torch.device] = None):
        self.config = dict(config or {})
        self.rank, self.world_size = _world()
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if device is None:
            device = torch.device("cuda", self.local_rank) if torch.cuda.is_available() else torch.device("cpu")
        self.device = device
        bf16 = bool(self.config.get("bf16", {}).get("enabled", False))
        model.to(device=device, dtype=torch.bfloat16 if bf16 else None)
        self.module = model
        self.optimizer = optimizer
        self.lr_scheduler = lr_scheduler
        self.gradient_accumulation_steps = int(self.config.get("gradient_accumulation_steps", 1))
        self.gradient_clipping = float(self.config.get("gradient_clipping", 0.0))
        zo = self.config.get("zero_optimization", {}) or {}
        self.bucket_elems = int(os.environ.get("RWKVTTS_BUCKET_ELEMS", 0)) or max(int(zo.get("reduce_bucket_size", 0)), 1 << 26)
        self.overlap_comm = os.environ.get("RWKVTTS_OVERLAP_COMM", "1") != "0"
        self.global_steps = 0
        self.micro_steps = 0
        self.comm_ms = None                    # filled by profile_comm()
        self._on_gpu = self.device.type == "cuda"
        self._nccl = self.world_size > 1 and dist.get_backend() == "nccl"
        self._comm_stream = torch.cuda.Stream(device=self.device) if (self._on_gpu and self.world_size > 1) else None
        # fused exchange over NVLink peer memory (csrc/adam.cu adam_p2p_kernel): one kernel per bucket slice does
        # reduce-scatter(AVG) + Adam + all-gather on symmetric-memory buffers, in-switch reduction / broadcast when the
        # fabric offers multicast.  Needs NCCL ranks on one node, bf16, no gradient clipping (the clip factor would need
        # the reduced norm before the update).  zero_optimization.fused_p2p / RWKVTTS_ZERO_P2P=1 turn it on.
        want = zo.get("fused_p2p", None)
        env = os.environ.get("RWKVTTS_ZERO_P2P")
        self._want_p2p = bool(self._nccl and self.gradient_clipping == 0 and bf16 and self.world_size <= 8
                              and (env == "1" or (env is None and want)))
        self._p2p = None
        self.exchange_events = None
        self._build_flat_buffers()
        self._install_hooks()

    # -- flat parameter / gradient space ---------------------------------------------------------
    def _build_flat_buffers(self):
        groups = self.optimizer.param_groups
        if len(groups) > 8:
            raise ValueError("at most 8 optimizer param groups")
        gid = {}
        for gi, g in enumerate(groups):
            g["params"] = [p for p in g["params"]]
            for p in g["params"]:
                gid[id(p)] = gi
        ordered, seen = [], set()
        for p in self.module.parameters():                       # model order ...
            if id(p) in gid and id(p) not in seen:
                ordered.append(p); seen.add(id(p))
        for g in groups:                                          # ... then anything the module does not own
            for p in g["params"]:
                if id(p) not in seen:
                    ordered.append(p); seen.add(id(p))
        if not ordered:
            raise ValueError("optimizer has no parameters")
        ordered.reverse()                                         # gradients arrive last layer first
        dtype = ordered[0].dtype
        if any(p.dtype != dtype for p in ordered):
            raise ValueError("all optimized parameters must share one dtype (cast the model first)")
        W = self.world_size
        unit = ALIGN * W
        self._params, self._poff, self._pgroup, self._pbucket = ordered, [], [], []
        self.buckets = []                                         # (start, end) in the flat space
        off, bstart = 0, 0
        for i, p in enumerate(ordered):
            self._poff.append(off)
            self._pgroup.append(gid[id(p)])
            self._pbucket.append(len(self.buckets))
            off += (p.numel() + 7) // 8 * 8
            if off - bstart >= self.bucket_elems or i == len(ordered) - 1:
                off = bstart + ((off - bstart + unit - 1) // unit) * unit
                self.buckets.append((bstart, off))
                bstart = off
        self.numel, self.padded = sum(p.numel() for p in ordered), off
        if self._want_p2p:
            self._p2p = self._symmetric_buffers(off, dtype)
        if self._p2p is not None:
            self.flat_param, self.flat_grad = self._p2p["param"], self._p2p["grad"]
            self.flat_param.zero_(); self.flat_grad.zero_()
        else:
            self.flat_param = torch.zeros(off, dtype=dtype, device=self.device)
            self.flat_grad = torch.zeros(off, dtype=dtype, device=self.device)
        for p, o in zip(ordered, self._poff):
            n = p.numel()
            self.flat_param[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[o:o + n].view(p.shape)          # parameters alias the flat buffer
            p.grad = self.flat_grad[o:o + n].view(p.shape)           # (what a reader of .grad sees; see _collect)
        self._gviews = [p.grad for p in ordered]
        self._pindex = {id(p): i for i, p in enumerate(ordered)}
        self._has = [False] * len(ordered)                           # this window's gradient is in the flat buffer
        self._stash = [[] for _ in self.buckets]                     # gradients autograd handed over, not yet in it
        # shard space: slice `rank` of every bucket, concatenated
        self._slice, self._soff = [], []
        so = 0
        for (s, e) in self.buckets:
            ss = (e - s) // W
            self._slice.append((s + self.rank * ss, s + (self.rank + 1) * ss))
            self._soff.append(so)
            so += ss
        self.shard_size = so
        self.shard = [tuple(x) for x in self._slice]                 # flat ranges this rank owns
        self.master = torch.cat([self.flat_param[a:b] for a, b in self._slice]).float()
        self.exp_avg = torch.zeros_like(self.master)
        self.exp_avg_sq = torch.zeros_like(self.master)
        self.grad_shard = torch.zeros(so, dtype=dtype, device=self.device) if W > 1 else None
        self._opt_step = [0] * len(groups)
        self._stat = torch.zeros(2, dtype=torch.float32, device=self.device)
        self._norm2 = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._skipped = torch.zeros(1, dtype=torch.int64, device=self.device)
        # segment tables of the rank's slices: exclusive ends (relative to the slice start) and param group ids
        self._segs = []
        for (a, b) in self._slice:
            ends, gids = [], []
            for o, p, gi in zip(self._poff, ordered, self._pgroup):
                pe = o + (p.numel() + 7) // 8 * 8
                lo, hi = max(o, a), min(pe, b)
                if lo < hi:
                    ends.append(hi - a); gids.append(gi)
            if not ends or ends[-1] < b - a:                     # bucket padding behind the last parameter
                ends.append(b - a); gids.append(gids[-1] if gids else 0)
            self._segs.append((torch.tensor(ends, dtype=torch.int64, device=self.device),
                               torch.tensor(gids, dtype=torch.int32, device=self.device), ends, gids))

    def _symmetric_buffers(self, numel, dtype):
        """Flat parameter and gradient buffers in symmetric memory, mapped into every rank of the node (and bound to a
        multicast object when the NVSwitch offers it).  Returns None -- on every rank -- if any rank could not set it up."""
        ok, res = 1, None
        try:
            import torch.distributed._symmetric_memory as symm
            bufs, hdls = {}, {}
            for name in ("param", "grad"):
                bufs[name] = symm.empty(numel, dtype=dtype, device=self.device)
                hdls[name] = symm.rendezvous(bufs[name], dist.group.WORLD)
            res = {"param": bufs["param"], "grad": bufs["grad"], "hp": hdls["param"], "hg": hdls["grad"],
                   "param_ptrs": [int(x) for x in hdls["param"].buffer_ptrs], "grad_ptrs": [int(x) for x in hdls["grad"].buffer_ptrs],
                   "mc_param": int(getattr(hdls["param"], "multicast_ptr", 0) or 0),
                   "mc_grad": int(getattr(hdls["grad"], "multicast_ptr", 0) or 0)}
            if os.environ.get("RWKVTTS_ZERO_MULTICAST") == "0":
                res["mc_param"] = res["mc_grad"] = 0
        except Exception as e:                           # no symmetric memory here: NCCL path
            ok = 0
            warnings.warn(f"symmetric memory unavailable ({e!r}); engine uses the NCCL exchange")
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return res if int(flag.item()) == 1 else None

    def _install_hooks(self):
        self._hook_handles = []
        self._pending = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._dirty = [False] * len(self.buckets)
        self._works = []
        self._bucket_np = [0] * len(self.buckets)
        for b in self._pbucket:
            self._bucket_np[b] += 1
        overlap = self._nccl and self.overlap_comm and self._p2p is None
        self._arrived = [0] * len(self.buckets)
        for i, (p, b) in enumerate(zip(self._params, self._pbucket)):
            def hook(param, i=i, b=b):
                # backward() cleared .grad, so autograd handed over the gradient tensor itself instead of adding it into
                # the flat buffer with a kernel per parameter: it is parked until its bucket is complete and then moved
                # with the bucket's other gradients in ONE launch (csrc/gather.cu::multi_copy_kernel)
                g = param.grad
                view = self._gviews[i]
                if g is not None and g.data_ptr() != view.data_ptr():
                    self._stash[b].append((i, g))
                    param.grad = view
                else:
                    self._has[i] = True                           # accumulated in place (somebody kept .grad bound)
                self._arrived[b] += 1
                if self._arrived[b] >= self._bucket_np[b]:
                    self._flush(b)
                if not overlap or not self.is_gradient_accumulation_boundary():
                    return
                if self._launched[b]:
                    self._dirty[b] = True                        # a gradient arrived after its bucket left: redo in step()
                    return
                self._pending[b] -= 1
                if self._pending[b] == 0:
                    self._flush(b)
                    self._launch_reduce(b)
            self._hook_handles.append(p.register_post_accumulate_grad_hook(hook))
        self._pending = list(self._bucket_np)

    def _flush(self, b=None):
        """Move the parked gradients of bucket b (all buckets if None) into the flat gradient buffer."""
        for bb in (range(len(self.buckets)) if b is None else (b,)):
            items = self._stash[bb]
            if not items:
                continue
            self._stash[bb] = []
            fast, dtype = [], self.flat_grad.dtype
            for i, g in items:
                if (self._on_gpu and g.dtype == dtype and g.is_contiguous() and g.data_ptr() % 16 == 0
                        and g.numel() == self._gviews[i].numel() and dtype in (torch.bfloat16, torch.float32)):
                    fast.append((i, g))
                elif self._has[i]:
                    self._gviews[i].add_(g.to(dtype).view_as(self._gviews[i]))
                else:
                    self._gviews[i].copy_(g.view_as(self._gviews[i]))
                    self._has[i] = True
            if fast:
                n = len(fast)
                srcs = (ctypes.c_void_p * n)(*[g.data_ptr() for _, g in fast])
                offs = (ctypes.c_longlong * n)(*[self._poff[i] for i, _ in fast])
                cnts = (ctypes.c_longlong * n)(*[g.numel() for _, g in fast])
                accs = (ctypes.c_int * n)(*[int(self._has[i]) for i, _ in fast])
                with torch.cuda.device(self.device):
                    rc = _lib.lib().rwkvtts_multi_copy(srcs, offs, cnts, accs, n, self.flat_grad.data_ptr(),
                                                       self.flat_grad.element_size(),
                                                       torch.cuda.current_stream(self.device).cuda_stream)
                _lib.check(rc, "rwkvtts_multi_copy")
                for i, _ in fast:
                    self._has[i] = True

    def _launch_reduce(self, b):
        s, e = self.buckets[b]
        so, ss = self._soff[b], (e - s) // self.world_size
        self._comm_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._comm_stream):
            w = dist.reduce_scatter_tensor(self.grad_shard[so:so + ss], self.flat_grad[s:e], op=dist.ReduceOp.AVG,
                                           async_op=True)
        self._works.append(w)
        self._launched[b] = True

    def _rebind_grads(self):
        for i, p in enumerate(self._params):
            g = self._gviews[i]
            if p.grad is None:
                p.grad = g
            elif p.grad.data_ptr() != g.data_ptr():      # something replaced .grad: fold it in
                g.add_(p.grad)
                self._has[i] = True
                p.grad = g

    def _zero_grads(self):
        self.flat_grad.zero_()
        self._has = [False] * len(self._params)

    # -- module surface ----------------------------------------------------------------------------
    def __call__(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    forward = __call__

    def __getattr__(self, name):             # only reached when normal lookup fails
        module = self.__dict__.get("module")
        if module is None:
            raise AttributeError(name)
        return getattr(module, name)

    def train(self, mode: bool = True):
        self.module.train(mode)
        return self

    def eval(self):
        self.module.eval()
        return self

    def parameters(self, recurse: bool = True):
        return self.module.parameters(recurse)

    def named_parameters(self, *a, **k):
        return self.module.named_parameters(*a, **k)

    def zero_grad(self, set_to_none: bool = False):
        self._stash = [[] for _ in self.buckets]
        self._zero_grads()
        self._rebind_grads()

    def get_lr(self):
        return [g["lr"] for g in self.optimizer.param_groups]

    def is_gradient_accumulation_boundary(self) -> bool:
        return (self.micro_steps + 1) % self.gradient_accumulation_steps == 0

    @property
    def global_grad_norm(self) -> float:
        """Gradient norm of the last step (reads a device scalar: a host sync only when somebody asks)."""
        if self._p2p is not None:
            return float(self._norm2[0].sqrt())
        return float(self._stat[0].sqrt())

    @property
    def skipped_steps(self) -> int:
        return int(self._skipped.item())

    # -- training step --------------------------------------------------------------------------
    def backward(self, loss: torch.Tensor, **kwargs):
        self._flush()
        self._rebind_grads()
        for p in self._params:
            p.grad = None                        # autograd then hands the gradient tensor to the hook (see _install_hooks)
        self._arrived = [0] * len(self.buckets)
        if self.gradient_accumulation_steps > 1:
            loss = loss / self.gradient_accumulation_steps
        loss.backward()
        self._flush()                            # buckets with parameters that got no gradient
        self._rebind_grads()
        return loss

    def _reduce_gradients(self):
        """Every bucket's gradients averaged over the ranks into grad_shard (world_size > 1)."""
        W = self.world_size
        if self._nccl:
            for b in range(len(self.buckets)):
                if not self._launched[b] or self._dirty[b]:      # parameters that got no gradient / a late gradient
                    self._launch_reduce(b)
            for w in self._works:
                w.wait()                                         # the compute stream waits for the exchange
            self._works = []
        else:           # gloo (CPU tests of the host logic) has no reduce-scatter
            dist.all_reduce(self.flat_grad)
            for (a, b), so in zip(self._slice, self._soff):
                self.grad_shard[so:so + b - a].copy_(self.flat_grad[a:b]).div_(W)
        self._pending = list(self._bucket_np)
        self._launched = [False] * len(self.buckets)
        self._dirty = [False] * len(self.buckets)

    def _grad_of(self, b):
        a, e = self._slice[b]
        if self.world_size > 1:
            return self.grad_shard[self._soff[b]:self._soff[b] + e - a]
        return self.flat_grad[a:e]

    def _group_hp(self):
        hp = []
        for gi, g in enumerate(self.optimizer.param_groups):
            b1, b2 = g["betas"]
            t = self._opt_step[gi]
            bc = g.get("bias_correction", True)
            hp.append((float(g["lr"]), float(g["weight_decay"]), 1.0 - b1 ** t if bc else 1.0,
                       math.sqrt(1.0 - b2 ** t) if bc else 1.0))
        return hp

    @torch.no_grad()
    def step(self):
        boundary = self.is_gradient_accumulation_boundary()
        self.micro_steps += 1
        if not boundary:
            return
        self._flush()
        self._rebind_grads()
        W = self.world_size
        if self._p2p is not None:
            return self._step_p2p()
        if W > 1:
            self._reduce_gradients()
        # global squared gradient norm and non-finite count: one small all-reduce, consumed on the device
        self._stat.zero_()
        if self._on_gpu:
            L = _lib.lib()
            st = torch.cuda.current_stream(self.device).cuda_stream
            with torch.cuda.device(self.device):
                if W > 1:
                    _lib.check(L.rwkvtts_grad_stat(self.grad_shard.data_ptr(), int(self.grad_shard.dtype == torch.bfloat16),
                                                   self.grad_shard.numel(), self._stat.data_ptr(), st), "rwkvtts_grad_stat")
                else:
                    _lib.check(L.rwkvtts_grad_stat(self.flat_grad.data_ptr(), int(self.flat_grad.dtype == torch.bfloat16),
                                                   self.flat_grad.numel(), self._stat.data_ptr(), st), "rwkvtts_grad_stat")
        else:
            gf = (self.grad_shard if W > 1 else self.flat_grad).float()
            self._stat.copy_(torch.stack([gf.pow(2).sum(), (~torch.isfinite(gf)).sum().float()]))
        if W > 1:
            dist.all_reduce(self._stat)
        for gi in range(len(self._opt_step)):
            self._opt_step[gi] += 1
        hp = self._group_hp()
        g0 = self.optimizer.param_groups[0]
        b1, b2 = g0["betas"]
        eps = float(g0["eps"])
        _invalidate_param_cache()        # parameters are rewritten below without any version counter moving
        gathers = []
        for b, (a, e) in enumerate(self._slice):
            so, n = self._soff[b], e - a
            g = self._grad_of(b)
            pslice = self.flat_param[a:e]
            if self._on_gpu:
                ends, gids, _, _ = self._segs[b]
                hp_c = (ctypes.c_float * (4 * len(hp)))(*[x for h in hp for x in h])
                with torch.cuda.device(self.device):
                    rc = _lib.lib().rwkvtts_adam_multi(
                        self.master[so:so + n].data_ptr(), self.exp_avg[so:so + n].data_ptr(),
                        self.exp_avg_sq[so:so + n].data_ptr(), g.data_ptr(), int(g.dtype == torch.bfloat16),
                        pslice.data_ptr(), int(pslice.dtype == torch.bfloat16), n, ends.data_ptr(), gids.data_ptr(),
                        ends.numel(), hp_c, len(hp), b1, b2, eps, self.optimizer.adam_w_mode, self._stat.data_ptr(),
                        self.gradient_clipping, self._skipped.data_ptr() if b == 0 else None,      # one count per step
                        torch.cuda.current_stream(self.device).cuda_stream)
                _lib.check(rc, "rwkvtts_adam_multi")
            else:
                self._cpu_adam(b, g, pslice, hp, b1, b2, eps)
            if W > 1:
                s, e2 = self.buckets[b]
                if self._nccl:
                    # in place (pslice is this rank's slice of the bucket); overlaps the next bucket's Adam kernel
                    self._comm_stream.wait_stream(torch.cuda.current_stream(self.device))
                    with torch.cuda.stream(self._comm_stream):
                        gathers.append(dist.all_gather_into_tensor(self.flat_param[s:e2], pslice, async_op=True))
                else:
                    dist.all_gather(list(self.flat_param[s:e2].chunk(W)), pslice.clone())
        for w in gathers:
            w.wait()
        _invalidate_param_cache()     # the all-gather rewrote the other ranks' slices of the flat parameter buffer
        self._zero_grads()
        self.global_steps += 1
        if self.lr_scheduler is not None:
            self.optimizer._opt_called = True     # the update ran above (on the shard): torch's scheduler order check looks here
            self.lr_scheduler.step()

    def _step_p2p(self):
        """The optimizer step with the exchange fused into the update kernel (see __init__ / csrc/adam.cu)."""
        P, L = self._p2p, _lib.lib()
        W = self.world_size
        st = torch.cuda.current_stream(self.device).cuda_stream
        # non-finite check on the LOCAL gradients (a non-finite value on any rank makes the sum non-finite): one pass over
        # the local buffer + a 2-float all-reduce; the fused kernels read the flag on the device
        self._stat.zero_()
        self._norm2.zero_()
        with torch.cuda.device(self.device):
            _lib.check(L.rwkvtts_grad_stat(self.flat_grad.data_ptr(), 1, self.flat_grad.numel(), self._stat.data_ptr(), st),
                       "rwkvtts_grad_stat")
        dist.all_reduce(self._stat)
        for gi in range(len(self._opt_step)):
            self._opt_step[gi] += 1
        hp = self._group_hp()
        g0 = self.optimizer.param_groups[0]
        b1, b2 = g0["betas"]
        eps = float(g0["eps"])
        hp_c = (ctypes.c_float * (4 * len(hp)))(*[x for h in hp for x in h])
        gp = (ctypes.c_void_p * W)(*P["grad_ptrs"])
        pp = (ctypes.c_void_p * W)(*P["param_ptrs"])
        _invalidate_param_cache()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        P["hg"].barrier(channel=0)                       # every rank's gradients are final and visible
        for b, (a, e) in enumerate(self._slice):
            so, n = self._soff[b], e - a
            ends, gids, _, _ = self._segs[b]
            with torch.cuda.device(self.device):
                rc = L.rwkvtts_adam_p2p(
                    self.master[so:so + n].data_ptr(), self.exp_avg[so:so + n].data_ptr(), self.exp_avg_sq[so:so + n].data_ptr(),
                    gp, pp, P["mc_grad"] or None, P["mc_param"] or None, W, a, n, ends.data_ptr(), gids.data_ptr(), ends.numel(),
                    hp_c, len(hp), b1, b2, eps, self.optimizer.adam_w_mode, self._stat.data_ptr(), self._norm2.data_ptr(),
                    self._skipped.data_ptr() if b == 0 else None, st)
            _lib.check(rc, "rwkvtts_adam_p2p")
        P["hp"].barrier(channel=0)                       # every rank's parameter writes have landed; gradients are free
        ev1.record()
        self.exchange_events = (ev0, ev1)
        dist.all_reduce(self._norm2)                     # reporting only (global_grad_norm); nothing waits for it
        self._zero_grads()
        self.global_steps += 1
        if self.lr_scheduler is not None:
            self.optimizer._opt_called = True
            self.lr_scheduler.step()

    def _cpu_adam(self, b, g, pslice, hp, b1, b2, eps):
        """Host-logic twin of rwkvtts_adam_multi (gloo / CPU tests): same segment walk, same skip and clipping rule."""
        norm = float(self._stat[0].sqrt())
        if float(self._stat[1]) > 0 or not math.isfinite(norm):
            if b == 0:
                self._skipped += 1
            return
        gscale = 1.0
        if self.gradient_clipping > 0 and norm > self.gradient_clipping:
            gscale = self.gradient_clipping / (norm + 1e-6)
        so = self._soff[b]
        _, _, ends, gids = self._segs[b]
        lo = 0
        for hi, gi in zip(ends, gids):
            sl = slice(so + lo, so + hi)
            lr, wd, bc1, bc2s = hp[gi]
            grp = {"lr": lr, "weight_decay": wd, "betas": (b1, b2), "eps": eps, "bias_correction": False}
            m, v, w = self.exp_avg[sl], self.exp_avg_sq[sl], self.master[sl]
            gg = g[lo:hi].float() * gscale
            if not self.optimizer.adam_w_mode:
                gg = gg + wd * w
            m.mul_(b1).add_(gg, alpha=1 - b1)
            v.mul_(b2).addcmul_(gg, gg, value=1 - b2)
            upd = (m / bc1) / (v.sqrt() / bc2s + eps)
            if self.optimizer.adam_w_mode:
                upd = upd + wd * w
            w.add_(upd, alpha=-lr)
            pslice[lo:hi].copy_(w)
            lo = hi

    @torch.no_grad()
    def profile_comm(self, iters: int = 5):
        """Device time of the step's exchange in isolation (all buckets: reduce-scatter, then all-gather), ms; the
        bench reports it next to the step time as the collective's share.  Leaves gradients zeroed."""
        if not self._nccl:
            self.comm_ms = {"reduce_scatter_ms": 0.0, "all_gather_ms": 0.0, "bytes_per_rank": 0}
            return self.comm_ms
        if self._p2p is not None:
            ms = None
            if self.exchange_events is not None:
                torch.cuda.synchronize(self.device)
                ms = self.exchange_events[0].elapsed_time(self.exchange_events[1])
            self.comm_ms = {"fused_exchange_and_adam_ms": ms, "reduce_scatter_ms": 0.0, "all_gather_ms": 0.0,
                            "bytes_per_rank": self.padded * self.flat_grad.element_size(), "buckets": len(self.buckets),
                            "mode": "one kernel per slice: reduce-scatter + Adam + all-gather over symmetric memory ("
                                    + ("NVSwitch multicast: multimem.ld_reduce / multimem.st" if self._p2p["mc_grad"] else
                                       "peer loads / stores") + ")"}
            return self.comm_ms
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        rs = ag = 0.0
        for _ in range(iters):
            torch.cuda.synchronize(self.device)
            ev[0].record()
            for b, (s, e) in enumerate(self.buckets):
                so, ss = self._soff[b], (e - s) // self.world_size
                dist.reduce_scatter_tensor(self.grad_shard[so:so + ss], self.flat_grad[s:e], op=dist.ReduceOp.AVG)
            ev[1].record()
            for b, (s, e) in enumerate(self.buckets):
                a, e2 = self._slice[b]
                dist.all_gather_into_tensor(self.flat_param[s:e], self.flat_param[a:e2])
            ev[2].record()
            torch.cuda.synchronize(self.device)
            rs += ev[0].elapsed_time(ev[1]); ag += ev[1].elapsed_time(ev[2])
        self._zero_grads()
        self.comm_ms = {"reduce_scatter_ms": rs / iters, "all_gather_ms": ag / iters,
                        "bytes_per_rank": self.padded * self.flat_grad.element_size(), "buckets": len(self.buckets)}
        return self.comm_ms

    # -- checkpoints (DeepSpeed directory layout) --------------------------------------------------
    def _layout(self):
        return {"numel": self.numel, "padded": self.padded, "buckets": [list(b) for b in self.buckets],
                "dp_world_size": self.world_size}

    def save_checkpoint(self, save_dir: str, tag: Optional[str] = None, client_state: Optional[dict] = None,
                        save_latest: bool = True):
        tag = tag if tag is not None else f"global_step{self.global_steps}"
        path = os.path.join(save_dir, str(tag))
        os.makedirs(path, exist_ok=True)
        if self.rank == 0:
            sd = {k: v.detach().cpu() for k, v in self.module.state_dict().items()}
            torch.save({"module": sd, "global_steps": self.global_steps, "micro_steps": self.micro_steps,
                        "skipped_steps": self.skipped_steps, "layout": self._layout(),
                        "dp_world_size": self.world_size,
                        "lr_scheduler": self.lr_scheduler.state_dict() if self.lr_scheduler is not None else None,
                        **(client_state or {})}, os.path.join(path, "mp_rank_00_model_states.pt"))
            if save_latest:
                with open(os.path.join(save_dir, "latest"), "w") as f:
                    f.write(str(tag))
        torch.save({"slices": [list(x) for x in self._slice], "layout": self._layout(), "master": self.master.cpu(),
                    "exp_avg": self.exp_avg.cpu(), "exp_avg_sq": self.exp_avg_sq.cpu(), "opt_step": self._opt_step,
                    "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.optimizer.param_groups]},
                   os.path.join(path, f"zero_pp_rank_{self.rank}_mp_rank_00_optim_states.pt"))
        if self.world_size > 1:
            dist.barrier()
        return True

    def _load_optimizer_states(self, path, saved_world):
        """Restores master / moments for this rank's slices.  Same layout: the rank's own file.  Different world size
        or bucket layout (same model): the saved ranks' slices are stitched into the full flat space and re-cut."""
        files = [os.path.join(path, f"zero_pp_rank_{r}_mp_rank_00_optim_states.pt") for r in range(saved_world)]
        mine = os.path.join(path, f"zero_pp_rank_{self.rank}_mp_rank_00_optim_states.pt")
        if os.path.exists(mine):
            os_ = torch.load(mine, map_location="cpu")
            if os_.get("layout") == self._layout() and [list(x) for x in self._slice] == os_.get("slices"):
                self.master.copy_(os_["master"]); self.exp_avg.copy_(os_["exp_avg"]); self.exp_avg_sq.copy_(os_["exp_avg_sq"])
                self._opt_step = list(os_["opt_step"])
                return True
        if not all(os.path.exists(f) for f in files):
            return False
        full = None
        for f in files:
            os_ = torch.load(f, map_location="cpu")
            if os_.get("layout", {}).get("numel") != self.numel:
                return False
            if full is None:
                full = {k: torch.zeros(max(os_["layout"]["padded"], self.padded)) for k in ("master", "exp_avg", "exp_avg_sq")}
            so = 0
            for (a, b) in os_["slices"]:
                for k in full:
                    full[k][a:b] = os_[k][so:so + b - a]
                so += b - a
            step = list(os_["opt_step"])
        # the flat order of parameters is fixed by the model; only the bucket padding can move offsets
        if os_["layout"]["buckets"] != [list(b) for b in self.buckets] and len(os_["layout"]["buckets"]) != 1 \
                and len(self.buckets) != 1:
            return False
        for (a, b), so in zip(self._slice, self._soff):
            self.master[so:so + b - a].copy_(full["master"][a:b])
            self.exp_avg[so:so + b - a].copy_(full["exp_avg"][a:b])
            self.exp_avg_sq[so:so + b - a].copy_(full["exp_avg_sq"][a:b])
        self._opt_step = step
        return True

    def load_checkpoint(self, load_dir: str, tag: Optional[str] = None, load_optimizer_states: bool = True, **kwargs):
        if tag is None:
            latest = os.path.join(load_dir, "latest")
            if not os.path.exists(latest):
                return None, None
            tag = open(latest).read().strip()
        path = os.path.join(load_dir, str(tag))
        ms = torch.load(os.path.join(path, "mp_rank_00_model_states.pt"), map_location="cpu")
        with torch.no_grad():
            own = self.module.state_dict()
            for k, v in ms["module"].items():
                own[k].copy_(v)                       # in place: parameters keep aliasing the flat buffer
        self.global_steps, self.micro_steps = ms["global_steps"], ms["micro_steps"]
        self._skipped.fill_(int(ms.get("skipped_steps", 0)))
        self.master.copy_(torch.cat([self.flat_param[a:b] for a, b in self._slice]).float())
        if load_optimizer_states:
            ok = False
            try:
                ok = self._load_optimizer_states(path, int(ms.get("dp_world_size", self.world_size)))
            except (OSError, KeyError, RuntimeError) as e:
                warnings.warn(f"optimizer states of {path} could not be read: {e!r}")
            if not ok:
                warnings.warn(f"load_checkpoint({path}): optimizer states were requested but could not be restored for this "
                              f"layout (saved world size {ms.get('dp_world_size')}, now {self.world_size}); Adam moments and "
                              "bias correction restart from zero, fp32 master weights are rebuilt from the model weights")
        if self.lr_scheduler is not None and ms.get("lr_scheduler") is not None:
            self.lr_scheduler.load_state_dict(ms["lr_scheduler"])
        client = {k: v for k, v in ms.items() if k not in ("module", "lr_scheduler")}
        return path, client


def initialize(args=None, model: torch.nn.Module = None, optimizer=None, model_parameters: Optional[Iterable] = None,
               training_data=None, lr_scheduler=None, config=None, config_params=None, **kwargs):
    """deepspeed.initialize(...) -> (engine, optimizer, training_dataloader, lr_scheduler)."""
    cfg = config if config is not None else config_params
    if isinstance(cfg, str):
        import json
        with open(cfg) as f:
            cfg = json.load(f)
    cfg = cfg or {}
    init_distributed(cfg.get("distributed_backend"))
    if model is None:
        raise ValueError("deepspeed.initialize: model is required")
    if optimizer is None:
        oc = cfg.get("optimizer", {}).get("params", {})
        params = list(model_parameters) if model_parameters is not None else list(model.parameters())
        optimizer = FusedAdam([p for p in params if p.requires_grad], lr=oc.get("lr", 1e-3),
                              betas=tuple(oc.get("betas", (0.9, 0.999))), eps=oc.get("eps", 1e-8),
                              weight_decay=oc.get("weight_decay", 0.0))
    elif not isinstance(optimizer, FusedAdam):
        raise TypeError("this engine drives rwkvtts_b200.engine.FusedAdam / DeepSpeedCPUAdam "
                        "(what `from deepspeed.ops.adam import ...` resolves to)")
    engine = Engine(model, optimizer, cfg, lr_scheduler=lr_scheduler)
    return engine, optimizer, None, lr_scheduler
