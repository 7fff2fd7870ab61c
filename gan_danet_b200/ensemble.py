"""Deep ensemble of GAN-DANet generators, one member per B200 (BASELINE.json configs[4]; SURVEY 8e "ensemble", 8f-f1).

Mirror of the reference's ``EnsembleTrainer`` (``/root/reference/deep_ensemble.ipynb:298-483``): same method names, seeds
(``42 + i``, :312), checkpoint names (``best_model_member_{i+1}.pth``, :338) and return values.  What changes is where the
work runs:

  * members are **sharded over the ranks** (member i lives on rank ``i % world``) instead of being trained / evaluated one
    after the other on one device (:323-340, :388-424) -- "replicas only", no data-path collective;
  * the monthly sweep keeps everything on the device: input preparation (``prepare_input_nhwc``), the generator in eval mode,
    inverse scaling (:415-416), plateau masking and spatial means (:450-460), mean / std over the members (:463-464) are
    kernels of ``libgandanet_sm100.so``; the reference concatenates numpy arrays on the host;
  * the only exchange is the final statistic: an ``all_gather`` of each rank's member fields (NCCL over NVLink; gloo in the
    CPU tests), padded with NaN members when the ensemble does not divide the world size -- the statistics kernels skip NaN
    exactly like ``np.nanmean`` / ``np.nanstd``.
"""
from __future__ import annotations

import copy
import os
import random
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import torch


def member_ranks(num_ensemble: int, world: int) -> List[List[int]]:
    """members_of[rank] = member indices (0-based) that live on that rank: round robin, member i on rank i % world."""
    return [[i for i in range(num_ensemble) if i % world == r] for r in range(world)]


def gather_members(local: torch.Tensor, num_ensemble: int, group=None) -> torch.Tensor:
    """[m_local, ...] member fields of this rank -> [num_ensemble, ...] in member order on every rank.

    One ``all_gather`` of equally sized blocks: ranks holding fewer members pad with NaN fields, which are dropped here."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert local.shape[0] == num_ensemble
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    owners = member_ranks(num_ensemble, world)
    assert local.shape[0] == len(owners[rank]), (local.shape, owners[rank])
    per = max(len(o) for o in owners)
    block = torch.full((per,) + tuple(local.shape[1:]), float("nan"), dtype=local.dtype, device=local.device)
    block[:local.shape[0]] = local
    parts = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(parts, block, group=group)
    out = torch.empty((num_ensemble,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r, members in enumerate(owners):
        for j, i in enumerate(members):
            out[i] = parts[r][j]
    return out


class EnsembleTrainer:
    """``EnsembleTrainer(num_ensemble, model_trainer_kwargs, ensemble_dir)`` of deep_ensemble.ipynb:298-313.

    ``model_trainer_kwargs`` is handed to ``trainer_factory`` (default: :func:`default_trainer_factory`), the stand-in for the
    notebook's ``ModelTrainer(**kwargs)`` whose data loading is outside the hot path."""

    def __init__(self, num_ensemble: int, model_trainer_kwargs: Optional[dict] = None, ensemble_dir: str = "ensemble_models",
                 trainer_factory: Optional[Callable[..., object]] = None, group=None):
        self.num_ensemble = num_ensemble
        self.model_trainer_kwargs = dict(model_trainer_kwargs or {})
        self.ensemble_dir = ensemble_dir
        os.makedirs(self.ensemble_dir, exist_ok=True)
        self.seeds = [42 + i for i in range(num_ensemble)]                    # :312
        self.trainer_factory = trainer_factory or default_trainer_factory
        self.group = group
        self._graphs: Dict[tuple, object] = {}

    # ------------------------------------------------------------------ placement
    def _world_rank(self) -> Tuple[int, int]:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group), dist.get_rank(self.group)
        return 1, 0

    def local_members(self) -> List[int]:
        world, rank = self._world_rank()
        return member_ranks(self.num_ensemble, world)[rank]

    def set_seed(self, seed: int) -> None:
        """:314-320."""
        import numpy as np
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(seed)

    def member_path(self, i: int) -> str:
        return os.path.join(self.ensemble_dir, f"best_model_member_{i + 1}.pth")

    # ------------------------------------------------------------------ training (:322-340)
    def train_ensemble(self, batches: Callable[[int], Iterable[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]], epochs: int = 1) -> Dict[int, List[float]]:
        """Trains the members that live on this rank, each from its own seed on the same data split (the notebook fixes
        ``rand = 42`` for the split, :328), and saves ``upsampling_module.state_dict()`` per member (:337-339).
        ``batches(epoch)`` yields device batches ``(lr_grace_05, lr_grace_025, hr_aux)``.  Returns the per-member loss_G curves."""
        curves: Dict[int, List[float]] = {}
        for i in self.local_members():
            self.set_seed(self.seeds[i])
            kwargs = copy.deepcopy(self.model_trainer_kwargs)
            kwargs["rand"] = 42
            trainer = self.trainer_factory(**kwargs)
            losses = []
            for epoch in range(epochs):
                for lr05, lr025, aux in batches(epoch):
                    losses.append(trainer.train_step(lr05, lr025, aux)["loss_G"])
                trainer.end_epoch()
            curves[i] = [float(v) for v in torch.stack(losses).cpu()] if losses else []
            torch.save(trainer.G.state_dict(), self.member_path(i))
        return curves

    # ------------------------------------------------------------------ loading (:342-366)
    def load_ensemble_models(self, model_class, device, input_channels: int, attention_type: Optional[str] = None, only_local: bool = True) -> list:
        """Loads this rank's members (all members when ``only_local`` is False or the job has one rank) in eval mode."""
        members = self.local_members() if only_local else list(range(self.num_ensemble))
        models = []
        for i in members:
            path = self.member_path(i)
            if not os.path.exists(path):
                raise FileNotFoundError(f"Model file '{path}' does not exist.")
            model = model_class(input_channels=input_channels, attention_type=attention_type).to(device)
            model.load_state_dict(torch.load(path, map_location=device))
            model.eval()
            models.append(model)
        return models

    # ------------------------------------------------------------------ monthly sweep (:368-428)
    def predict_ensemble(self, models: Sequence[torch.nn.Module], test_loader: Iterable, scaler: Tuple[float, float] = (1.0, 0.0), use_graph: bool = False):
        """All months through every local member.  ``test_loader`` yields ``(lr_grace_05, lr_grace_025, hr_aux)`` batches
        (host or device); ``scaler = (scale_, mean_)`` of the GRACE ``StandardScaler`` (:415-416).  ``use_graph`` replays one
        captured CUDA graph per (member, batch shape) instead of enqueueing the ~170 launches of each forward (inference.py).

        Returns ``(all_preds [m_local, T, 1, H, W], all_trues [T, 1, H, W])`` as de-standardised **device** tensors."""
        from . import postprocess as PP
        from .inference import GraphedGenerator
        from .trainer import generator_forward_nhwc, prepare_input_nhwc
        batches = list(test_loader)
        preds, trues = [], []
        with torch.no_grad():
            for idx, model in enumerate(models):
                dev = next(model.parameters()).device
                per_model = []
                for lr05, lr025, aux in batches:
                    lr05, aux = lr05.to(dev, non_blocking=True), aux.to(dev, non_blocking=True)
                    if use_graph:
                        key = (id(model), tuple(lr05.shape), tuple(aux.shape))
                        if key not in self._graphs:                       # captured once per (member, batch shape), replayed by every later sweep
                            self._graphs[key] = GraphedGenerator(model, lr05, aux)
                        per_model.append(self._graphs[key](lr05, aux))
                    else:
                        x = prepare_input_nhwc(lr05, aux)                                                         # :389-397
                        per_model.append(generator_forward_nhwc(model, x))                                        # :400
                    if idx == 0:
                        trues.append(lr025.to(dev, non_blocking=True))
                preds.append(PP.destandardise(torch.cat(per_model, dim=0), scaler[0], scaler[1]))
        all_trues = PP.destandardise(torch.cat(trues, dim=0), scaler[0], scaler[1]) if trues else None
        return torch.stack(preds, dim=0), all_trues

    # ------------------------------------------------------------------ statistics (:430-473)
    def spatial_means(self, fields: torch.Tensor, keep: Optional[torch.Tensor]) -> torch.Tensor:
        from . import postprocess as PP
        return PP.masked_spatial_mean(fields, keep)

    def compute_uncertainty(self, all_preds: torch.Tensor, trues: torch.Tensor, keep: Optional[torch.Tensor] = None):
        """``all_preds`` [m_local, T, C, H, W] (this rank's members), ``trues`` [T, C, H, W], ``keep`` [H, W] = ``tpb_h != 0``.
        Returns ``(mean_preds [T, C], std_preds [T, C], r2)`` as numpy / float like the notebook.  The member time series
        (tiny) are gathered over the ranks; the fields themselves never leave their GPU."""
        from . import postprocess as PP
        preds_ts = gather_members(self.spatial_means(all_preds, keep), self.num_ensemble, self.group)     # [M, T, C]
        trues_ts = self.spatial_means(trues, keep)                                                          # [T, C]
        mean_preds, std_preds = PP.ensemble_stats(preds_ts)
        t, m = trues_ts.double().cpu(), mean_preds.double().cpu()
        ok = ~torch.isnan(t) & ~torch.isnan(m)
        tv, mv = t[ok], m[ok]
        r2 = float(1.0 - ((tv - mv) ** 2).sum() / ((tv - tv.mean()) ** 2).sum())                             # sklearn r2_score (:471)
        return mean_preds.cpu().numpy(), std_preds.cpu().numpy(), r2

    def pixel_statistics(self, all_preds: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Per-pixel ensemble mean and spread [T, C, H, W] over ALL members (fields gathered over NVLink)."""
        from . import postprocess as PP
        return PP.ensemble_stats(gather_members(all_preds, self.num_ensemble, self.group))

    def save_uncertainty(self, std_preds, save_path: str = "ensemble_uncertainty.npy") -> None:
        import numpy as np
        np.save(save_path, std_preds)

    def save_mean_predictions(self, mean_preds, save_path: str = "ensemble_mean_predictions.npy") -> None:
        import numpy as np
        np.save(save_path, mean_preds)


def default_trainer_factory(input_channels: int = 46, attention_type: Optional[str] = "danet", epochs: int = 150, device=None, perceptual: bool = True,
                            sample_hw: Tuple[int, int] = (256, 512), rand: int = 42, **_ignored):
    """The notebook's ``ModelTrainer.__init__`` (GAN_DANet_train.ipynb:146-194) without its file loading: G, D, VGG19 perceptual
    term (random init offline), initialised under the seed that is current when it is called."""
    from . import FlexibleUpsamplingModule, Discriminator1, PerceptualLoss
    from .trainer import GANTrainer, init_like_reference
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    G = FlexibleUpsamplingModule(input_channels, attention_type=attention_type)
    D = Discriminator1()
    init_like_reference(G, D, torch.zeros(1, 1, *sample_hw))
    G, D = G.to(dev), D.to(dev)
    perc = None
    if perceptual:
        perc = PerceptualLoss(pretrained=False, device=torch.device("cpu"))
        perc.vgg.to(dev)
        perc.device = dev
    return GANTrainer(G, D, perc, epochs=epochs)
