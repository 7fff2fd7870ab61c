"""north_star: "a 200-step loss trajectory within 1 %" -- teacher-forced form (SURVEY 8c addendum, criterion (i)).

A free-running GAN trajectory is not reproducible to 1 % over 200 steps even by the reference against itself in another
precision (SURVEY 8c addendum: loss_D differs by 5 % at step 20 between the reference's fp32 and fp64 runs;
profiles/r01_trajectory_envelope.txt measures the same growth here), so the 200 steps are walked by the CPU oracle
(float32 -- the reference's own CPU arithmetic) with the faithful schedule (w = epoch/epochs, CosineAnnealingWarmRestarts per
epoch, four batches per epoch, GAN_DANet_train.ipynb:186-187,225-295), and AT EVERY STEP the CUDA trainer starts from the
oracle's full state (G and D parameters, BatchNorm buffers, AdamW moments and step count), runs the same step and must give
loss_D and loss_G within 1 % and post-step parameters within 2e-3.  Four modes are driven from the same oracle walk: the fp32 parity engine, the
tensor-core parity mode, the BENCHMARKED mode (bench.py) and the all-bf16 mode it replaced.

Measured on a B200 (profiles/r02_trajectory_teacher_forced.json), worst deviation over the 200 steps (loss_D / loss_G / parameters):
  fp32      fp32 engine                                                            3.1e-6 / 3.1e-4 (one step; median 3.6e-7) / 2.6e-5
  bf16x3    hi+lo split convolutions + fused PAM with split logits (parity mode)    2.4e-4 / 6.7e-4 / 1.1e-3
  bf16+gx3  BENCHMARKED: bf16 operands, the FORWARD convolutions of G (engine.generator_forward_x3) and of D on hi+lo split operands, every
            gradient GEMM and VGG19 on single bf16 operands                         2.4e-4 / 6.8e-4 / 1.1e-3
  bf16      bf16 operands in G's forward too (bench.py --g-forward bf16)            1.2e-2 / 1.1e-2 (median 1.2e-3 / 1.6e-3; 3 of 400 values above 1 %) / 3.5e-3
Asserted: fp32 engine, parity mode AND the benchmarked mode 1 % at EVERY step (north_star's bar); the all-bf16 mode 1 % on at least 95 % of the steps
and 2 % at every step, parameters 5e-3 (its generated field is 9e-3 from the reference's -- the cost of bf16 forward operands for ANY implementation:
tools/trajectory_quantised_cpu.py walks the same steps on the CPU with the reference algorithm and bf16 operands and finds loss_D above 1 % on 7 steps).
The per-step deviations are written to gpurun_out/r02_trajectory_teacher_forced.json when that directory exists.
A second test walks two teacher-forced steps at the NORTH-STAR grid (64x128: PAM over N = 8192 positions, 64 key tiles per query tile).
"""
import json
import os

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STEPS = int(os.environ.get("GDN_TRAJ_STEPS", "200"))
WATCH_G = ("final.weight", "upsample.0.weight", "dense_blocks.1.layers.2.conv.weight", "attention_modules.0.fuse.0.weight")
WATCH_D = ("fc2.weight", "conv1.weight", "conv3.weight")


@pytest.fixture(scope="module", autouse=True)
def _require_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gan_danet_b200 import _lib
    _lib.lib_for_device(0)


def _load_state(tr, st):
    """Oracle state -> CUDA trainer: parameters and buffers in place (optimizer keeps its parameter objects), AdamW moments."""
    tr.G.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in st.g.items()})
    tr.D.load_state_dict({k: v.float() for k, v in st.d.items()})
    for opt, mod, m, v in ((tr.opt_G, tr.G, st.m_g, st.v_g), (tr.opt_D, tr.D, st.m_d, st.v_d)):
        for name, p in mod.named_parameters():
            s = opt.state[p]
            if not s:
                assert st.step == 0, "optimizer state is created by the first step"
                continue
            s["step"] = st.step if isinstance(s["step"], int) else torch.tensor(float(st.step))
            s["exp_avg"].copy_(m[name])
            s["exp_avg_sq"].copy_(v[name])


def test_teacher_forced_trajectory_200_steps(oracle):
    import gan_danet_b200 as P
    from gan_danet_b200 import engine as E
    from gan_danet_b200.synthetic import make_batch
    from gan_danet_b200.trainer import GANTrainer
    epochs, per_epoch = 150, 4
    batches = [make_batch(10 * i, 2, 8, 16) for i in range(per_epoch)]
    torch.manual_seed(11)
    G0 = P.FlexibleUpsamplingModule(46)
    D0 = P.Discriminator1()
    G0.apply(P.weights_init_normal)
    for mod in (D0.conv1, D0.conv2, D0.conv3, D0.conv4, D0.fc2):
        mod.apply(P.weights_init_normal)
    D0._materialise_fc1(batches[0][1])
    with torch.no_grad():
        for n, p in G0.named_parameters():
            if n.endswith("gamma"):
                p.fill_(0.05)              # attention path live from the first step (faithful init 0 would keep PAM/CAM out of the loss)
    torch.manual_seed(12)
    vgg_sd = {k: v.clone() for k, v in P.PerceptualLoss(pretrained=False, device=torch.device("cpu")).vgg.state_dict().items()}
    st = oracle.TrainState({k: v.clone() for k, v in G0.state_dict().items()}, {k: v.clone() for k, v in D0.state_dict().items()}, vgg_sd)

    # conv precision, PAM precision, loss tol, parameter tol.  'bf16x3' = the tensor-core PARITY mode (hi+lo split conv operands + fused PAM with split
    # logits): 1 % at EVERY step, like the fp32 engine.  'bf16' = the all-bf16 mode (bench.py --g-forward bf16): its generated field is 1e-2 away from the reference's
    # (bf16 operands), which reaches D's logits -- 1 % on >= 95 % of the steps, 2 % everywhere (measured: 3 of 400 loss values above 1 %, worst 1.2 %)
    # 'bf16+gx3' = the BENCHMARKED mode (bench.py): bf16 operands everywhere except the forward convolutions of G (engine.generator_forward_x3) and of D,
    # which run on hi+lo split operands -- the generated field is then the parity mode's (1.5e-4 from the reference) and the losses must be within 1 % at
    # EVERY step and the post-step parameters within 2e-3, like the fp32 engine and the parity mode; every gradient GEMM still reads single bf16 operands
    modes = {"fp32": ("fp32", "fp32", 1e-2, 2e-3), "bf16x3": ("bf16x3", "fp16x3", 1e-2, 2e-3), "bf16": ("bf16", "fp16x3", 1e-2, 5e-3),
             "bf16+gx3": ("bf16", "fp16x3", 1e-2, 2e-3)}
    trainers = {}
    for name, (conv, pam, _, _) in modes.items():
        import copy
        G, D = copy.deepcopy(G0).to(DEV), copy.deepcopy(D0).to(DEV)
        perc = P.PerceptualLoss(pretrained=False, device=torch.device("cpu"))
        perc.vgg.load_state_dict(vgg_sd)
        perc.vgg.to(DEV)
        perc.device = torch.device(DEV)
        G.set_pam_precision(pam)
        tr = GANTrainer(G, D, perc, epochs=epochs)
        tr._ensure_opt_D(batches[0][1].to(DEV))
        trainers[name] = tr
    dev_batches = [tuple(t.to(DEV) for t in b) for b in batches]

    log = {name: [] for name in modes}
    worst = {name: {"loss_D": 0.0, "loss_G": 0.0, "param": 0.0, "update": 0.0} for name in modes}
    failures = []
    old, old_gx3 = E.conv_precision, E.generator_forward_x3
    try:
        for i in range(STEPS):
            epoch = i // per_epoch
            lr_d, lr_g = oracle.cosine_warm_restarts_lr(epoch, 4e-4), oracle.cosine_warm_restarts_lr(epoch, 2e-4)
            before = {"g": {k: st.g[k].clone() for k in WATCH_G}, "d": {k: st.d[k].clone() for k in WATCH_D}}
            for name, tr in trainers.items():
                _load_state(tr, st)
            ref = oracle.train_step(st, *batches[i % per_epoch], epoch=epoch, epochs=epochs, lr_d=lr_d, lr_g=lr_g)     # st is now the post-step state
            for name, tr in trainers.items():
                conv, _, tol_l, tol_p = modes[name]
                E.set_conv_precision(conv)
                E.generator_forward_x3 = name == "bf16+gx3"
                tr.epoch = epoch
                for opt, lr in ((tr.opt_D, lr_d), (tr.opt_G, lr_g)):
                    for grp in opt.param_groups:
                        grp["lr"] = lr
                out = tr.train_step(*dev_batches[i % per_epoch])
                rec = {"step": i}
                for k in ("loss_D", "loss_G"):
                    dev = abs(float(out[k]) - ref[k]) / max(abs(ref[k]), 1e-3)
                    rec[k] = dev
                    worst[name][k] = max(worst[name][k], dev)
                    if dev > (2 * tol_l if name == "bf16" else tol_l):
                        failures.append((name, i, k, float(out[k]), ref[k]))
                pe, ue = 0.0, 0.0
                for mod, ref_sd, b4, names in ((tr.G, st.g, before["g"], WATCH_G), (tr.D, st.d, before["d"], WATCH_D)):
                    sd = mod.state_dict()
                    for k in names:
                        pe = max(pe, rel_err(sd[k], ref_sd[k]))
                        ue = max(ue, rel_err(sd[k].cpu().double() - b4[k].double(), ref_sd[k].double() - b4[k].double()))
                rec["param"], rec["update"] = pe, ue
                worst[name]["param"], worst[name]["update"] = max(worst[name]["param"], pe), max(worst[name]["update"], ue)
                if pe > tol_p:
                    failures.append((name, i, "param", pe, tol_p))
                log[name].append(rec)
        torch.cuda.synchronize()
    finally:
        E.set_conv_precision(old)
        E.generator_forward_x3 = old_gx3
        out_dir = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(out_dir):
            with open(os.path.join(out_dir, "r02_trajectory_teacher_forced.json"), "w") as f:
                json.dump({"steps": len(log["fp32"]), "worst": worst, "final_ref": ref if STEPS else None, "log": log}, f)
    print("teacher-forced worst deviations:", json.dumps(worst))
    assert not failures, failures[:8]
    for k in ("loss_D", "loss_G"):
        over = sum(r[k] > 1e-2 for r in log["bf16"])
        assert over <= 0.05 * max(len(log["bf16"]), 1), (k, over)



def test_teacher_forced_steps_at_the_north_star_grid(oracle):
    """Two teacher-forced G+D steps at grid 64x128 (output 256x512, PAM over N = 8192: the fused kernels' full key loop, lazy-reference rescaling, the
    padded d / C tiles), batch 1, against the CPU oracle in float32 -- every mode, losses within 1 % (all-bf16 mode 2 %); measured in the benchmarked mode 'bf16+gx3': <= 2.8e-4."""
    import copy
    import gan_danet_b200 as P
    from gan_danet_b200 import engine as E
    from gan_danet_b200.synthetic import fast_batch
    from gan_danet_b200.trainer import GANTrainer
    epochs = 150
    batches = [fast_batch(500 + i, 1, 64, 128) for i in range(2)]
    torch.manual_seed(21)
    G0 = P.FlexibleUpsamplingModule(46)
    D0 = P.Discriminator1()
    G0.apply(P.weights_init_normal)
    D0.apply(P.weights_init_normal)
    D0._materialise_fc1(batches[0][1])
    with torch.no_grad():
        for n, p in G0.named_parameters():
            if n.endswith("gamma"):
                p.fill_(0.05)
    torch.manual_seed(22)
    vgg_sd = {k: v.clone() for k, v in P.PerceptualLoss(pretrained=False, device=torch.device("cpu")).vgg.state_dict().items()}
    st = oracle.TrainState({k: v.clone() for k, v in G0.state_dict().items()}, {k: v.clone() for k, v in D0.state_dict().items()}, vgg_sd)
    modes = {"fp32": ("fp32", "fp32", 1e-2), "bf16x3": ("bf16x3", "fp16x3", 1e-2), "bf16": ("bf16", "fp16x3", 2e-2), "bf16+gx3": ("bf16", "fp16x3", 1e-2)}
    trainers = {}
    for name, (conv, pam, _) in modes.items():
        G, D = copy.deepcopy(G0).to(DEV), copy.deepcopy(D0).to(DEV)
        perc = P.PerceptualLoss(pretrained=False, device=torch.device("cpu"))
        perc.vgg.load_state_dict(vgg_sd)
        perc.vgg.to(DEV)
        perc.device = torch.device(DEV)
        G.set_pam_precision(pam)
        tr = GANTrainer(G, D, perc, epochs=epochs)
        tr.epoch = 3
        tr._ensure_opt_D(batches[0][1].to(DEV))
        trainers[name] = tr
    old, old_gx3 = E.conv_precision, E.generator_forward_x3
    report = {}
    try:
        for i, b in enumerate(batches):
            for tr in trainers.values():
                _load_state(tr, st)
            ref = oracle.train_step(st, *b, epoch=3, epochs=epochs)
            for name, tr in trainers.items():
                conv, _, tol = modes[name]
                E.set_conv_precision(conv)
                E.generator_forward_x3 = name == "bf16+gx3"
                out = tr.train_step(*(t.to(DEV) for t in b))
                for k in ("loss_D", "loss_G", "pixel", "perceptual"):
                    dev = abs(float(out[k]) - ref[k]) / max(abs(ref[k]), 1e-3)
                    report[(name, i, k)] = dev
                    assert dev <= tol, (name, i, k, float(out[k]), ref[k])
    finally:
        E.set_conv_precision(old)
        E.generator_forward_x3 = old_gx3
        E.release_buffers()
    print("north-star-grid teacher-forced deviations:", {f"{k[0]}/{k[1]}/{k[2]}": round(v, 6) for k, v in report.items()})
