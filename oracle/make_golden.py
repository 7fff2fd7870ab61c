"""Generates the golden fixtures under ``tests/golden/`` by running the UNMODIFIED reference
(``/root/reference/models`` imported read-only) on seeded inputs, in float64 on the CPU.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The fixtures pin ``oracle/gan_danet_oracle.py`` (tests/test_oracle_golden.py) -- the reference itself ships no tests or
golden vectors for this path.  Values are stored as float32 (inputs are float32-exact).
"""
from __future__ import annotations

import os
import sys

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def ref_models():
    sys.path.insert(0, "/root/reference")
    import importlib
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    gen = importlib.import_module("models.generator")
    m = importlib.import_module("models")
    sys.path.pop(0)
    return m, gen


def f32(t):
    return t.detach().to(torch.float32).contiguous()


def save(name, obj):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".pt")
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def randn(shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


def module_case(mod, x, r):
    """Run ``mod`` in float64: y, d(sum(y*r))/dx and parameter gradients."""
    mod = mod.double()
    xd = x.double().requires_grad_(True)
    y = mod(xd)
    (y * r.double()).sum().backward()
    return {"sd": {k: f32(v) for k, v in mod.state_dict().items()}, "x": x, "r": r, "y": f32(y), "dx": f32(xd.grad),
            "grads": {k: f32(p.grad) for k, p in mod.named_parameters() if p.grad is not None}}


def srgand_case(M):
    """J: SRGAND (models/discriminator.py:8-54; exported, never instantiated by the notebooks): dim 8, input [2,1,128,128] -> 2x2 before the pool."""
    torch.manual_seed(5)
    D = M.SRGAND(dim=8, in_channels=1)
    D.apply(M.weights_init_normal)
    D = D.double().train()
    x = randn((2, 1, 128, 128), 60)
    xg = x.double().requires_grad_(True)
    z = D(xg)
    (z * torch.tensor([[1.0], [-0.5]], dtype=torch.float64)).sum().backward()
    save("srgand_dim8_128x128", {"seed": 5, "dim": 8, "x": x, "logits": f32(z), "dx": f32(xg.grad),
                                 "grads_small": {k: f32(p.grad) for k, p in D.named_parameters() if p.numel() <= 30000},
                                 "grad_norms": {k: float(p.grad.norm()) for k, p in D.named_parameters()},
                                 "buffers_after": {k: f32(v) for k, v in D.state_dict().items() if "running" in k},
                                 "keys": list(D.state_dict().keys())})


def main():
    if "--only-srgand" in sys.argv:
        M, _ = ref_models()
        srgand_case(M)
        return
    M, gen = ref_models()
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---- A/B/C: attention modules, gamma = 0.5 so the attention path is live (SURVEY 8c caveat 4)
    for C, (h, w) in ((160, (8, 16)), (184, (4, 8))):
        torch.manual_seed(0)
        pam = gen.PAMModule(C)
        pam.apply(M.weights_init_normal)
        with torch.no_grad():
            pam.gamma.fill_(0.5)
            pam.query.bias.copy_(0.1 * randn((C // 8,), 11))      # non-zero biases so the bias paths are exercised
            pam.key.bias.copy_(0.1 * randn((C // 8,), 12))
            pam.value.bias.copy_(0.1 * randn((C,), 13))
        x = 0.5 * randn((2, C, h, w), 1)
        r = randn((2, C, h, w), 2)
        save(f"pam_c{C}_{h}x{w}", module_case(pam, x, r))
        cam = gen.CAMModule(C)
        with torch.no_grad():
            cam.gamma.fill_(0.5)
        xc = 0.12 * randn((2, C, h, w), 3)          # keeps softmax(-E) away from one-hot so gradients are informative
        save(f"cam_c{C}_{h}x{w}", module_case(cam, xc, r))
    torch.manual_seed(0)
    da = gen.DANetAttention(160)
    da.apply(M.weights_init_normal)
    with torch.no_grad():
        da.position_attention.gamma.fill_(0.5)
        da.channel_attention.gamma.fill_(0.5)
    da.train()
    case = module_case(da, 0.3 * randn((2, 160, 8, 16), 4), randn((2, 160, 8, 16), 5))
    del case["sd"]["fuse.0.weight"], case["grads"]["fuse.0.weight"]     # 1.8 MB each; regenerated from the seed by the test
    save("danet_c160_8x16", case)

    # ---- D: dense block + transition
    torch.manual_seed(0)
    db = gen.DenseBlock(4, 64, 24)
    db.apply(M.weights_init_normal)
    db.train()
    save("denseblock_64_8x16", module_case(db, randn((2, 64, 8, 16), 6), randn((2, 160, 8, 16), 7)))
    torch.manual_seed(0)
    tl = gen.TransitionLayer(160, 80)
    tl.apply(M.weights_init_normal)
    tl.train()
    save("transition_160_8x16", module_case(tl, randn((2, 160, 8, 16), 8), randn((2, 80, 8, 16), 9)))

    # ---- E: whole generator (C_in 46, grid 8x16, B 2), gamma = 0.05; weights come from the seed
    torch.manual_seed(0)
    G = M.FlexibleUpsamplingModule(46)
    G.apply(M.weights_init_normal)
    with torch.no_grad():
        for m in G.modules():
            if isinstance(m, (gen.PAMModule, gen.CAMModule)):
                m.gamma.fill_(0.05)
    G = G.double().train()
    x = randn((2, 46, 8, 16), 20)
    r = randn((2, 1, 32, 64), 21)
    xd = x.double().requires_grad_(True)
    y = G(xd)
    (y * r.double()).sum().backward()
    small = {k: f32(p.grad) for k, p in G.named_parameters() if p.numel() <= 30000}
    norms = {k: float(p.grad.norm()) for k, p in G.named_parameters()}
    bufs = {k: f32(v) for k, v in G.state_dict().items() if "running" in k and ("initial" in k or "upsample.5" in k or "fuse" in k)}
    save("generator_cin46_8x16", {"seed": 0, "gamma": 0.05, "x": x, "r": r, "y": f32(y), "dx": f32(xd.grad), "grads_small": small,
                                  "grad_norms": norms, "buffers_after": bufs, "keys": list(G.state_dict().keys()),
                                  "n_params": sum(p.numel() for p in G.parameters())})

    # ---- F: discriminator on [2,1,64,128]; weights from the seed (fc1: default nn.Linear init after a dummy forward)
    torch.manual_seed(3)
    D = M.Discriminator1()
    xd0 = randn((2, 1, 64, 128), 30)
    for mod in (D.conv1, D.conv2, D.conv3, D.conv4, D.fc2):
        mod.apply(M.weights_init_normal)
    with torch.no_grad():
        D(xd0)                                   # materialise fc1 (consumes the RNG for its default init)
    D = D.double()
    xg = xd0.double().requires_grad_(True)
    z = D(xg)
    (z * torch.tensor([[1.0], [-0.5]], dtype=torch.float64)).sum().backward()
    save("discriminator_64x128", {"seed": 3, "x": xd0, "logits": f32(z), "dx": f32(xg.grad),
                                  "grads_small": {k: f32(p.grad) for k, p in D.named_parameters() if p.numel() <= 30000},
                                  "grad_norms": {k: float(p.grad.norm()) for k, p in D.named_parameters()},
                                  "keys": list(D.state_dict().keys())})

    # ---- G: losses on [2,1,32,64]
    a, b = randn((2, 1, 32, 64), 40), randn((2, 1, 32, 64), 41)
    ad = a.double().requires_grad_(True)
    tv = M.TVLoss(1e-5)(ad)
    (dtv,) = torch.autograd.grad(tv, ad)
    ssim = M.SSIM().double()(a.double(), b.double())
    torch.manual_seed(2)
    perc = M.PerceptualLoss(pretrained=False, device=torch.device("cpu"))
    perc.vgg.double()
    pl = perc(ad, b.double())
    (dpl,) = torch.autograd.grad(pl, ad)
    zz = randn((6, 1), 42)
    bce1 = F.binary_cross_entropy_with_logits(zz.double(), torch.ones(6, 1, dtype=torch.float64))
    bce0 = F.binary_cross_entropy_with_logits(zz.double(), torch.zeros(6, 1, dtype=torch.float64))
    save("losses_32x64", {"a": a, "b": b, "tv": float(tv), "dtv": f32(dtv), "ssim": float(ssim), "perceptual": float(pl), "dperceptual": f32(dpl),
                          "vgg_seed": 2, "z": zz, "bce1": float(bce1), "bce0": float(bce0), "mse": float(((a.double() - b.double()) ** 2).mean())})

    # ---- H: resampling semantics of ATen (what the reference calls)
    t = randn((1, 3, 5, 7), 50).double()
    save("resample", {"x": f32(t), "up2": f32(F.interpolate(t, scale_factor=2, mode="bicubic", align_corners=False)),
                      "bil": f32(F.interpolate(t, size=(20, 28), mode="bilinear", align_corners=False)),
                      "x8": f32(randn((1, 2, 8, 12), 51)),
                      "down2": f32(F.interpolate(randn((1, 2, 8, 12), 51).double(), scale_factor=0.5, mode="bicubic")),
                      "down4": f32(F.interpolate(randn((1, 2, 8, 12), 51).double(), scale_factor=0.25, mode="bicubic"))})

    # ---- I: two full training steps with the reference modules + torch.optim.AdamW (the notebook's loop, :225-269)
    sys.path.insert(0, ROOT)
    from gan_danet_b200.synthetic import make_batch
    lr05, real, aux = make_batch(0, 2, 8, 16)
    torch.manual_seed(0)
    G = M.FlexibleUpsamplingModule(46)
    D = M.Discriminator1()
    G.apply(M.weights_init_normal)
    for mod in (D.conv1, D.conv2, D.conv3, D.conv4, D.fc2):
        mod.apply(M.weights_init_normal)
    with torch.no_grad():
        D(real)
    torch.manual_seed(2)
    P = M.PerceptualLoss(pretrained=False, device=torch.device("cpu"))
    G, D = G.double(), D.double()
    P.vgg.double()
    opt_D = torch.optim.AdamW(D.parameters(), lr=4e-4, betas=(0.5, 0.999), weight_decay=1e-4)
    opt_G = torch.optim.AdamW(G.parameters(), lr=2e-4, betas=(0.5, 0.999), weight_decay=1e-4)
    bce, mse, tvl, ssim_m = torch.nn.BCEWithLogitsLoss(), torch.nn.MSELoss(), M.TVLoss(1e-5), M.SSIM().double()
    hist = []
    epochs, epoch = 150, 3
    G.train()
    D.train()
    for _ in range(2):
        lr_grace = F.interpolate(lr05.double(), scale_factor=0.5, mode="bicubic")
        daux = F.interpolate(aux.double(), scale_factor=0.25, mode="bicubic")
        xin = torch.cat([lr_grace, daux], dim=1)
        hr = G(xin)
        opt_D.zero_grad()
        real_o, fake_o = D(real.double()), D(hr.detach())
        loss_D = (bce(real_o, torch.ones_like(real_o)) + bce(fake_o, torch.zeros_like(fake_o))) / 2
        loss_D.backward()
        opt_D.step()
        opt_G.zero_grad()
        fake_o = D(hr)
        adv, pix = bce(fake_o, torch.ones_like(fake_o)), mse(hr, real.double())
        ss = 1 - ssim_m(hr, real.double())
        tv_, pe = tvl(hr), P(hr, real.double())
        w = epoch / epochs
        loss_G = (1 - w) * pix + w * adv + tv_ + pe
        loss_G.backward()
        opt_G.step()
        hist.append({"loss_D": float(loss_D), "loss_G": float(loss_G), "adv": float(adv), "pixel": float(pix), "ssim": float(ss),
                     "tv": float(tv_), "perceptual": float(pe)})
    srgand_case(M)
    save("train_2steps_8x16", {"seed": 0, "vgg_seed": 2, "epoch": epoch, "epochs": epochs, "history": hist,
                               "final_w": f32(G.final.weight), "d_fc2_w": f32(D.fc2.weight), "initial_bn_rm": f32(G.initial[1].running_mean)})


if __name__ == "__main__":
    main()
