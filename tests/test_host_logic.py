"""Host-side logic that needs no GPU: schedules, API mirror, data-parallel gradient exchange (gloo, world_size 2)."""
import math
import os
import socket

import pytest
import torch


def test_cosine_warm_restarts_matches_torch():
    """CosineAnnealingWarmRestarts(T_0=10, T_mult=2, eta_min=1e-6) stepped per epoch (GAN_DANet_train.ipynb:186-187)."""
    from gan_danet_b200.trainer import cosine_warm_restarts_lr
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=2e-4)
    sch = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=10, T_mult=2, eta_min=1e-6)
    for epoch in range(0, 75):
        assert math.isclose(opt.param_groups[0]["lr"], cosine_warm_restarts_lr(epoch, 2e-4), rel_tol=1e-9, abs_tol=1e-15), epoch
        opt.step()
        sch.step()


def test_api_mirror_of_reference_models():
    """Same exports, constructor signatures and state_dict keys as models/__init__.py:12-23, generator.py:178-185."""
    import inspect
    import gan_danet_b200.models as M
    assert set(M.__all__) == {"CBAMBlock", "FlexibleUpsamplingModule", "OriginalRelationshipLearner", "SqueezeExcitation", "Discriminator1",
                              "SRGAND", "PerceptualLoss", "SSIM", "TVLoss", "weights_init_normal"}
    sig = inspect.signature(M.FlexibleUpsamplingModule.__init__)
    assert [(k, v.default) for k, v in list(sig.parameters.items())[1:]] == [
        ("input_channels", 40), ("growth_rate", 24), ("num_blocks", 3), ("num_layers_per_block", 4), ("attention_type", "danet")]
    G = M.FlexibleUpsamplingModule()
    assert len(G.state_dict()) == 163 and sum(p.numel() for p in G.parameters()) == 2268537
    assert G.feature_channels == [160, 176, 184]
    with pytest.warns(RuntimeWarning):
        G2 = M.FlexibleUpsamplingModule(attention_type="senet")       # aliases to DANet (generator.py:166-171)
    assert list(G2.state_dict().keys()) == list(G.state_dict().keys())
    assert M.FlexibleUpsamplingModule(attention_type=None).attention_modules[0] is None
    with pytest.raises(ValueError):
        M.FlexibleUpsamplingModule(attention_type="bogus")
    D = M.Discriminator1()
    D.apply(M.weights_init_normal)                                      # lazy fc1 is skipped, as on the authors' torch
    D._materialise_fc1(torch.zeros(1, 1, 180, 88))
    assert D.fc1.weight.shape == (1024, 512 * 12 * 6) and sum(p.numel() for p in D.parameters()) == 39300609
    assert list(M.SRGAND().state_dict().keys())[0] == "conv1.weight"
    M.PerceptualLoss(pretrained=False, use_gpu=False)                   # notebook passes use_gpu= (GAN_DANet_train.ipynb:194)


def test_pitch_of_slices():
    from gan_danet_b200.engine import pitch_of, rows_of
    buf = torch.zeros(2, 3, 5, 16)
    assert pitch_of(buf) == 16 and pitch_of(buf[..., 4:8]) == 16 and rows_of(buf[..., 4:8]) == 30
    assert pitch_of(torch.zeros(4, 1, 1, 7)) == 7
    assert pitch_of(torch.zeros(1, 9, 1, 4)) == 4
    with pytest.raises(Exception):
        pitch_of(buf.permute(0, 2, 1, 3))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gan_danet_b200.trainer import GradientAllReduce
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 2_000_000, 17, 3)]      # one "big" tensor, three bucketed ones
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float((rank + 1) * (i + 1)))
    params.append(torch.nn.Parameter(torch.zeros(4)))                                 # no grad: must be skipped
    ar = GradientAllReduce()
    ar(params)
    ok = all(torch.all(p.grad == float(sum(r + 1 for r in range(world)) * (i + 1))) for i, p in enumerate(params[:4]))
    out[rank] = bool(ok) and ar.world == world
    dist.destroy_process_group()


def _shard_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gan_danet_b200.trainer import GradientAllReduce
    g = torch.Generator().manual_seed(5)
    p0 = [torch.randn(n, generator=g) for n in (7, 2_000_000, 33, 1_500_001)]        # a shardable big tensor, small ones, a big one NOT divisible by the world size
    gr = [[torch.randn(p.shape, generator=g) for p in p0] for _ in range(world)]     # every rank knows every rank's gradient (for the expected value)
    lr = 0.1

    def run(shard):
        params = [torch.nn.Parameter(p.clone()) for p in p0]
        for p, gg in zip(params, gr[rank]):
            p.grad = gg.clone()
        ar = GradientAllReduce(shard_big=shard)
        ar.start(params)()
        with torch.no_grad():
            for p in params:          # a plain SGD step on whatever slice this rank owns (the optimizer-side contract of shard_of)
                r = ar.shard_of(p)
                if r is None:
                    p -= lr * p.grad
                else:
                    p.view(-1)[r[0]:r[1]] -= lr * p.grad.view(-1)[r[0]:r[1]]
        ar.gather_params(params)
        return ar, params

    ar_s, sharded = run(True)
    _, plain = run(False)
    want = [p - lr * sum(gr[r][i] for r in range(world)) for i, p in enumerate(p0)]
    ok = ar_s.shard_of(sharded[1]) == (rank * 1_000_000, (rank + 1) * 1_000_000) and ar_s.shard_of(sharded[0]) is None and ar_s.shard_of(sharded[3]) is None
    for a, b, w in zip(sharded, plain, want):
        ok = ok and torch.allclose(a, w, rtol=1e-6, atol=1e-6) and torch.allclose(b, w, rtol=1e-6, atol=1e-6)
    # the updated parameters are identical on every rank
    chk = torch.stack([a.detach().double().sum() for a in sharded])
    all_chk = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(all_chk, chk)
    ok = ok and all(torch.equal(c, all_chk[0]) for c in all_chk)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_big_tensor_gloo_world2():
    """GradientAllReduce(shard_big=True): reduce-scatter of the big gradient, slice-wise update, all-gather of the parameter == all-reduce + full update
    (Discriminator1.fc1 under data parallelism); tensors that do not divide by the world size keep the all-reduce."""
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_shard_worker, args=(world, port, out), nprocs=world, join=True)
        assert out[0] and out[1]


def test_gradient_allreduce_gloo_world2():
    """The N > 1 path: gradients are summed over ranks (bucketed small tensors + in-place large ones)."""
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_dp_worker, args=(world, port, out), nprocs=world, join=True)
        assert out[0] and out[1]


# ------------------------------------------------------------------------------------------------ ensemble sharding
def test_member_placement():
    from gan_danet_b200.ensemble import EnsembleTrainer, member_ranks
    assert member_ranks(8, 8) == [[i] for i in range(8)]                        # BASELINE configs[4]: one member per B200
    assert member_ranks(5, 2) == [[0, 2, 4], [1, 3]]
    assert member_ranks(3, 1) == [[0, 1, 2]]
    import tempfile
    ens = EnsembleTrainer(5, {"epochs": 2}, ensemble_dir=tempfile.mkdtemp())
    assert ens.seeds == [42, 43, 44, 45, 46]                                    # deep_ensemble.ipynb:312
    assert ens.member_path(0).endswith("best_model_member_1.pth")               # :338
    assert ens.local_members() == [0, 1, 2, 3, 4]
    with pytest.raises(FileNotFoundError):
        ens.load_ensemble_models(None, "cpu", 46)


def _ens_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gan_danet_b200.ensemble import EnsembleTrainer, gather_members
    import tempfile
    M = 5
    ens = EnsembleTrainer(M, {}, ensemble_dir=tempfile.mkdtemp())
    mine = ens.local_members()
    local = torch.stack([torch.full((2, 3), float(i)) for i in mine])          # member i's "field" is the constant i
    full = gather_members(local, M)
    ok = full.shape == (M, 2, 3) and all(bool((full[i] == float(i)).all()) for i in range(M)) and not torch.isnan(full).any()
    out[rank] = bool(ok) and mine == ([0, 2, 4] if rank == 0 else [1, 3])
    dist.destroy_process_group()


def test_ensemble_gather_gloo_world2():
    """Uneven ensemble (5 members on 2 ranks): NaN-padded all_gather, members back in index order on every rank."""
    import torch.multiprocessing as mp
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_ens_worker, args=(2, port, out), nprocs=2, join=True)
        assert out[0] and out[1]


def test_generator_forward_precision_scope():
    """engine.generator_forward_x3: inside the product mode ('bf16') the generator records its forward under conv_precision_scope('bf16x3'); in every
    other mode the flag does nothing, and the scope restores the outer mode (also when the body raises)."""
    from gan_danet_b200 import engine as E
    old, old_flag = E.conv_precision, E.generator_forward_x3
    try:
        E.generator_forward_x3 = True
        for mode, want in (("bf16", "bf16x3"), ("bf16x3", None), ("fp32", None)):
            E.set_conv_precision(mode)
            assert E.generator_forward_precision() == want
            with E.conv_precision_scope(E.generator_forward_precision()):
                assert E.conv_precision == (want or mode)
                assert not (want and E.bf16_storage_ok())          # the bf16-only feature storage is off while the forward is recorded with split operands
                assert E.backward_conv_precision() == mode          # ... and the backward closures will run under the outer mode
                assert E.split_forward_in_product_mode() == (want is not None)
            assert E.conv_precision == mode and E.backward_conv_precision() == mode and not E.split_forward_in_product_mode()
        E.set_conv_precision("bf16")
        try:
            with E.conv_precision_scope(E.generator_forward_precision()):
                raise RuntimeError("body failed")
        except RuntimeError:
            pass
        assert E.conv_precision == "bf16"
        E.generator_forward_x3 = False
        assert E.generator_forward_precision() is None
    finally:
        E.set_conv_precision(old)
        E.generator_forward_x3 = old_flag
