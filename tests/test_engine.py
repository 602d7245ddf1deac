"""Host logic of the DeepSpeed-compatible engine on CPU: single process, and world_size 2 over gloo
(SURVEY.md section 8e).  The sharded reduce-scatter -> Adam -> all-gather step must equal a plain
full-batch AdamW on one process."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _model(seed=0):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(24, 40), torch.nn.Tanh(), torch.nn.Linear(40, 8))


def _groups(m):
    dec = [p for n, p in m.named_parameters() if n.endswith("weight")]
    nod = [p for n, p in m.named_parameters() if not n.endswith("weight")]
    return [{"params": nod, "weight_decay": 0.0, "my_lr_scale": 2.0, "name": "lr_2x"},
            {"params": dec, "weight_decay": 0.1, "my_lr_scale": 1.0, "name": "lr_decay"}]


def _data(n=16, seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 24, generator=g), torch.randn(n, 8, generator=g)


def _reference_steps(steps, lr=1e-2):
    m = _model()
    opt = torch.optim.AdamW([{"params": g["params"], "weight_decay": g["weight_decay"]} for g in _groups(m)],
                            lr=lr, betas=(0.9, 0.95), eps=1e-18)
    x, y = _data()
    for _ in range(steps):
        opt.zero_grad()
        torch.nn.functional.mse_loss(m(x), y).backward()
        opt.step()
    return [p.detach().clone() for p in m.parameters()]


def _engine_steps(steps, rank=0, world=1, lr=1e-2, gas=1):
    import deepspeed
    from deepspeed.ops.adam import FusedAdam
    m = _model()
    opt = FusedAdam(_groups(m), lr=lr, betas=(0.9, 0.95), eps=1e-18, bias_correction=True, adam_w_mode=True,
                    amsgrad=False, weight_decay=0.1)
    eng, opt2, _, _ = deepspeed.initialize(model=m, config={"train_batch_size": 16, "bf16": {"enabled": False},
                                                            "gradient_accumulation_steps": gas,
                                                            "zero_optimization": {"stage": 2}},
                                           model_parameters=m.parameters(), optimizer=opt)
    assert opt2 is opt and eng.module is m
    x, y = _data()
    n = x.shape[0] // world
    xs, ys = x[rank * n:(rank + 1) * n], y[rank * n:(rank + 1) * n]      # DistributedSampler-style shard
    for _ in range(steps):
        for mb in range(gas):
            k = n // gas
            loss = torch.nn.functional.mse_loss(eng(xs[mb * k:(mb + 1) * k]), ys[mb * k:(mb + 1) * k])
            eng.backward(loss)
            eng.step()
    return eng


def test_single_process_matches_adamw():
    eng = _engine_steps(5)
    for a, b in zip(eng.parameters(), _reference_steps(5)):
        assert torch.allclose(a, b, atol=2e-6), (a - b).abs().max()
    assert eng.global_steps == 5 and eng.local_rank == 0


def test_gradient_accumulation_matches_full_batch():
    eng = _engine_steps(3, gas=2)
    for a, b in zip(eng.parameters(), _reference_steps(3)):
        assert torch.allclose(a, b, atol=2e-6)
    assert eng.global_steps == 3 and eng.micro_steps == 6


def test_attribute_passthrough_and_lr_mutation():
    eng = _engine_steps(1)
    assert eng.training is True
    eng.eval()
    assert eng.module.training is False
    for g in eng.optimizer.param_groups:            # the scripts overwrite lr every step
        g["lr"] = 0.0
    before = [p.detach().clone() for p in eng.parameters()]
    x, y = _data()
    eng.train()
    eng.backward(torch.nn.functional.mse_loss(eng(x), y))
    eng.step()
    for a, b in zip(eng.parameters(), before):
        assert torch.equal(a, b)


def test_nonfinite_gradient_skips_the_step():
    eng = _engine_steps(1)
    before = [p.detach().clone() for p in eng.parameters()]
    x, y = _data()
    loss = torch.nn.functional.mse_loss(eng(x), y) * float("nan")
    eng.backward(loss)
    eng.step()
    assert eng.skipped_steps == 1 and eng.global_steps == 2      # DeepSpeed counts a skipped step as a step
    for a, b in zip(eng.parameters(), before):
        assert torch.equal(a, b)


def test_checkpoint_round_trip():
    eng = _engine_steps(2)
    with tempfile.TemporaryDirectory() as d:
        eng.save_checkpoint(d)
        assert open(os.path.join(d, "latest")).read() == "global_step2"
        assert os.path.exists(os.path.join(d, "global_step2", "mp_rank_00_model_states.pt"))
        assert os.path.exists(os.path.join(d, "global_step2", "zero_pp_rank_0_mp_rank_00_optim_states.pt"))
        eng2 = _engine_steps(0)
        path, client = eng2.load_checkpoint(d)
        assert path.endswith("global_step2") and eng2.global_steps == 2
    x, y = _data()
    for e in (eng, eng2):
        e.backward(torch.nn.functional.mse_loss(e(x), y))
        e.step()
    for a, b in zip(eng.parameters(), eng2.parameters()):
        assert torch.allclose(a, b, atol=1e-7)


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(1)
    eng = _engine_steps(4, rank=rank, world=world)
    assert eng.world_size == world and eng.shard[0][0] == eng.buckets[0][0] + rank * eng.shard_size
    if rank == 0:
        torch.save([p.detach().clone() for p in eng.parameters()], out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_gloo_matches_single_process():
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "p.pt")
        mp.spawn(_worker, args=(2, 29541 + os.getpid() % 200, out), nprocs=2, join=True)
        got = torch.load(out)
    for a, b in zip(got, _reference_steps(4)):
        assert torch.allclose(a, b, atol=2e-6), (a - b).abs().max()


def test_optimizer_step_invalidates_the_no_grad_parameter_cache():
    """The Adam kernel writes parameters through raw pointers and the engine's parameters alias a flat buffer, so an
    optimizer step moves no parameter's autograd version counter; the cached fp32 / stacked copies that the no_grad
    paths (prefill, decode) keep must be dropped by the step itself."""
    import deepspeed
    from deepspeed.ops.adam import FusedAdam
    from rwkvtts_b200 import fused
    m = _model()
    opt = FusedAdam(_groups(m), lr=1e-2, betas=(0.9, 0.95), eps=1e-18)
    eng, _, _, _ = deepspeed.initialize(model=m, config={"train_batch_size": 16, "bf16": {"enabled": False},
                                                         "zero_optimization": {"stage": 2}},
                                        model_parameters=m.parameters(), optimizer=opt)
    w = m[0].weight
    calls = []
    def copy():
        calls.append(1)
        return w.detach().double().clone()
    with torch.no_grad():
        c0 = fused.cached(w, (w,), "test", copy)
        assert fused.cached(w, (w,), "test", copy) is c0 and len(calls) == 1          # hit
    ver = w._version
    x, y = _data()
    eng.backward(torch.nn.functional.mse_loss(eng(x), y))
    eng.step()
    assert w._version == ver                       # the hazard: nothing autograd-visible happened to the parameter
    with torch.no_grad():
        c1 = fused.cached(w, (w,), "test", copy)
    assert len(calls) == 2 and not torch.equal(c0, c1) and torch.equal(c1, w.detach().double())
    # in-place edits that do move the version counter are still seen without a step
    with torch.no_grad():
        w.mul_(2.0)
        c2 = fused.cached(w, (w,), "test", copy)
    assert len(calls) == 3 and torch.equal(c2, 2.0 * c1)


def test_multiple_buckets_match_adamw(monkeypatch):
    """Small buckets: the toy model is cut into several; slices, segment tables and the shard space must still give
    exactly AdamW."""
    monkeypatch.setenv("RWKVTTS_BUCKET_ELEMS", "300")
    eng = _engine_steps(5)
    assert len(eng.buckets) >= 2
    for a, b in zip(eng.parameters(), _reference_steps(5)):
        assert torch.allclose(a, b, atol=2e-6), (a - b).abs().max()


def test_gradient_clipping_matches_torch():
    import deepspeed
    from deepspeed.ops.adam import FusedAdam
    m = _model()
    opt = FusedAdam(_groups(m), lr=1e-2, betas=(0.9, 0.95), eps=1e-18)
    eng, _, _, _ = deepspeed.initialize(model=m, config={"bf16": {"enabled": False}, "gradient_clipping": 0.05},
                                        model_parameters=m.parameters(), optimizer=opt)
    ref = _model()
    ropt = torch.optim.AdamW([{"params": g["params"], "weight_decay": g["weight_decay"]} for g in _groups(ref)],
                             lr=1e-2, betas=(0.9, 0.95), eps=1e-18)
    x, y = _data()
    for _ in range(3):
        eng.backward(torch.nn.functional.mse_loss(eng(x), y)); eng.step()
        ropt.zero_grad(); torch.nn.functional.mse_loss(ref(x), y).backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.05); ropt.step()
    assert eng.global_grad_norm > 0.05
    for a, b in zip(eng.parameters(), ref.parameters()):
        assert torch.allclose(a, b, atol=5e-6), (a - b).abs().max()


def _ckpt_worker(rank, world, port, d):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(1)
    eng = _engine_steps(3, rank=rank, world=world)
    eng.save_checkpoint(d)
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_checkpoint_saved_on_two_ranks_resumes_on_one(recwarn):
    """Re-sharding: the two ranks' optimizer slices are stitched back and re-cut for world size 1; the next step equals
    the uninterrupted single-process run (a resume on another GPU count silently restarted Adam)."""
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_ckpt_worker, args=(2, 29741 + os.getpid() % 200, d), nprocs=2, join=True)
        eng = _engine_steps(0)
        eng.load_checkpoint(d)
        assert not [w for w in recwarn.list if "could not be restored" in str(w.message)]
        x, y = _data()
        eng.backward(torch.nn.functional.mse_loss(eng(x), y)); eng.step()
    for a, b in zip(eng.parameters(), _reference_steps(4)):
        assert torch.allclose(a, b, atol=3e-6), (a - b).abs().max()


def test_missing_optimizer_state_warns():
    eng = _engine_steps(2)
    with tempfile.TemporaryDirectory() as d:
        eng.save_checkpoint(d)
        os.remove(os.path.join(d, "global_step2", "zero_pp_rank_0_mp_rank_00_optim_states.pt"))
        eng2 = _engine_steps(0)
        with pytest.warns(UserWarning, match="could not be restored"):
            eng2.load_checkpoint(d)
