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

import math
import os
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
        self.global_steps = 0
        self.micro_steps = 0
        self.skipped_steps = 0
        self._build_flat_buffers()

    # -- flat parameter / gradient space ---------------------------------------------------------
    def _build_flat_buffers(self):
        groups = self.optimizer.param_groups
        plist: List[torch.nn.Parameter] = []
        self._segments = []            # (start, end, group index) in the flat space
        off = 0
        for gi, g in enumerate(groups):
            g["params"] = [p for p in g["params"]]
            start = off
            for p in g["params"]:
                plist.append(p)
                off += p.numel()
            self._segments.append((start, off, gi))
        if not plist:
            raise ValueError("optimizer has no parameters")
        dtype = plist[0].dtype
        if any(p.dtype != dtype for p in plist):
            raise ValueError("all optimized parameters must share one dtype (cast the model first)")
        unit = ALIGN * self.world_size
        total = ((off + unit - 1) // unit) * unit
        self.numel, self.padded = off, total
        self.flat_param = torch.zeros(total, dtype=dtype, device=self.device)
        self.flat_grad = torch.zeros(total, dtype=dtype, device=self.device)
        o = 0
        for p in plist:
            n = p.numel()
            self.flat_param[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[o:o + n].view(p.shape)          # parameters alias the flat buffer
            p.grad = self.flat_grad[o:o + n].view(p.shape)           # autograd accumulates in place
            o += n
        self._params = plist
        self.shard_size = total // self.world_size
        lo = self.rank * self.shard_size
        self.shard = (lo, lo + self.shard_size)
        self.master = self.flat_param[lo:lo + self.shard_size].float().clone()
        self.exp_avg = torch.zeros_like(self.master)
        self.exp_avg_sq = torch.zeros_like(self.master)
        self.grad_shard = torch.zeros(self.shard_size, dtype=dtype, device=self.device) if self.world_size > 1 else None
        self._opt_step = [0] * len(groups)

    def _rebind_grads(self):
        o = 0
        for p in self._params:
            n = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.flat_grad[o:o + n].data_ptr():
                g = self.flat_grad[o:o + n].view(p.shape)
                if p.grad is not None:                   # something replaced .grad (e.g. zero_grad(set_to_none))
                    g.add_(p.grad)
                p.grad = g
            o += n

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
        self.flat_grad.zero_()
        self._rebind_grads()

    def get_lr(self):
        return [g["lr"] for g in self.optimizer.param_groups]

    def is_gradient_accumulation_boundary(self) -> bool:
        return (self.micro_steps + 1) % self.gradient_accumulation_steps == 0

    # -- training step --------------------------------------------------------------------------
    def backward(self, loss: torch.Tensor, **kwargs):
        self._rebind_grads()
        if self.gradient_accumulation_steps > 1:
            loss = loss / self.gradient_accumulation_steps
        loss.backward()
        return loss

    @torch.no_grad()
    def step(self):
        boundary = self.is_gradient_accumulation_boundary()
        self.micro_steps += 1
        if not boundary:
            return
        self._rebind_grads()
        lo, hi = self.shard
        if self.world_size > 1:
            if dist.get_backend() == "nccl":
                dist.reduce_scatter_tensor(self.grad_shard, self.flat_grad, op=dist.ReduceOp.AVG)
            else:       # gloo (CPU tests of the host logic) has no reduce-scatter
                dist.all_reduce(self.flat_grad)
                self.grad_shard.copy_(self.flat_grad[lo:hi]).div_(self.world_size)
            g = self.grad_shard
        else:
            g = self.flat_grad[lo:hi]
        # global gradient norm and non-finite check in one small all-reduce
        gf = g.float()
        stat = torch.stack([gf.pow(2).sum(), (~torch.isfinite(gf)).any().float()])
        if self.world_size > 1:
            dist.all_reduce(stat)
        self.global_grad_norm = float(stat[0].sqrt())
        if float(stat[1]) > 0 or not math.isfinite(self.global_grad_norm):
            self.skipped_steps += 1              # same decision on every rank
            self.flat_grad.zero_()
            return
        gscale = 1.0
        if self.gradient_clipping > 0 and self.global_grad_norm > self.gradient_clipping:
            gscale = self.gradient_clipping / (self.global_grad_norm + 1e-6)
        pshard = self.flat_param[lo:hi]
        for gi in range(len(self._opt_step)):
            self._opt_step[gi] += 1
        for (s, e, gi) in self._segments:
            a, b = max(s, lo), min(e, hi)
            if a >= b:
                continue
            sl = slice(a - lo, b - lo)
            adam_update(self.master[sl], self.exp_avg[sl], self.exp_avg_sq[sl], g[sl], pshard[sl],
                        self.optimizer.param_groups[gi], self._opt_step[gi], self.optimizer.adam_w_mode, gscale)
        if self.world_size > 1:
            if dist.get_backend() == "nccl":
                dist.all_gather_into_tensor(self.flat_param, pshard)          # in place: pshard is rank's slice
            else:
                dist.all_gather(list(self.flat_param.chunk(self.world_size)), pshard.clone())
        _invalidate_param_cache()     # the all-gather rewrote the other ranks' shards of the flat parameter buffer
        self.flat_grad.zero_()
        self.global_steps += 1
        if self.lr_scheduler is not None:
            self.optimizer._opt_called = True     # the update ran above (on the shard): torch's scheduler order check looks here
            self.lr_scheduler.step()

    # -- checkpoints (DeepSpeed directory layout) --------------------------------------------------
    def save_checkpoint(self, save_dir: str, tag: Optional[str] = None, client_state: Optional[dict] = None,
                        save_latest: bool = True):
        tag = tag if tag is not None else f"global_step{self.global_steps}"
        path = os.path.join(save_dir, str(tag))
        os.makedirs(path, exist_ok=True)
        if self.rank == 0:
            sd = {k: v.detach().cpu() for k, v in self.module.state_dict().items()}
            torch.save({"module": sd, "global_steps": self.global_steps, "micro_steps": self.micro_steps,
                        "skipped_steps": self.skipped_steps, "dp_world_size": self.world_size,
                        "lr_scheduler": self.lr_scheduler.state_dict() if self.lr_scheduler is not None else None,
                        **(client_state or {})}, os.path.join(path, "mp_rank_00_model_states.pt"))
            if save_latest:
                with open(os.path.join(save_dir, "latest"), "w") as f:
                    f.write(str(tag))
        torch.save({"shard": self.shard, "padded": self.padded, "master": self.master.cpu(),
                    "exp_avg": self.exp_avg.cpu(), "exp_avg_sq": self.exp_avg_sq.cpu(), "opt_step": self._opt_step,
                    "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.optimizer.param_groups]},
                   os.path.join(path, f"zero_pp_rank_{self.rank}_mp_rank_00_optim_states.pt"))
        if self.world_size > 1:
            dist.barrier()
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
        self.skipped_steps = ms.get("skipped_steps", 0)
        lo, hi = self.shard
        self.master.copy_(self.flat_param[lo:hi].float())
        opt_file = os.path.join(path, f"zero_pp_rank_{self.rank}_mp_rank_00_optim_states.pt")
        if load_optimizer_states and os.path.exists(opt_file) and ms.get("dp_world_size") == self.world_size:
            os_ = torch.load(opt_file, map_location="cpu")
            if tuple(os_["shard"]) == tuple(self.shard):
                self.master.copy_(os_["master"]); self.exp_avg.copy_(os_["exp_avg"]); self.exp_avg_sq.copy_(os_["exp_avg_sq"])
                self._opt_step = list(os_["opt_step"])
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
