"""Training-side entry points of the hot path — host-side mirror of the pieces of the reference's `utils.py` that
`train_ovssc.py` / `train_vool.py` call: `config_parser` (utils.py:35-136, same flags and
defaults), `seed_all` (:216-222), `get_net` (:237-294: network + LAMB + HF scheduler + checkpoint resume with the
`module.` prefix rule), `compute_grad_norm` (:321-327), `loop` (:383-471: one pass over a loader, train or eval branch,
cross-rank stat averaging) and `train` (:474-614: epochs, `latest.pth` / `ckpt_<epoch>.pth`), plus re-exports of the
device metrics (`prediction_analysis`, `voxelize_points`, `get_bce_weight`).

What differs on purpose:
  * no `DistributedDataParallel` wrapper: gradients are averaged by `train.all_reduce_gradients` (NCCL all-reduce with
    DDP's find_unused_parameters semantics) between backward and the optimiser sweep.  Checkpoints written while
    torch.distributed is initialised still carry DDP's `module.` key prefix, so they interchange with the reference's;
    loading accepts both spellings;
  * `--use_amp` selects the single-fp16-operand mode of the convolution kernels (`precise=False`) instead of torch
    autocast + GradScaler (gradients are scaled per tensor on the device, unet3d_bwd.py); `scaler` is always None;
  * the optimiser is `train.Lamb` (one fused multi-tensor sweep with the clip folded in), numerically the reference's
    `arm.optim.lamb.Lamb` (tests/test_train_gpu.py).
The dataset classes (dataset.py: HDF5 scenes) are outside the hot path: `train()` takes any mapping-style datasets that
yield the batch contract of SURVEY.md §8d; `SyntheticOVSSCDataset` / `SyntheticVOOLDataset` produce it from seeds.
"""
from __future__ import annotations

import logging
import os
import random
from argparse import ArgumentParser
from typing import Callable, Dict, Optional

import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

from . import train as _train
from .clip.wrapper import saliency_configs
from .metrics import prediction_analysis, voxel_prediction_analysis, voxelize_points  # noqa: F401  (re-exports)
from .train import Lamb, get_bce_weight  # noqa: F401

# flag table of utils.config_parser (utils.py:35-136): (name, type, default) — store_true flags keep the reference's
# defaults (several default to True there, which makes them constants; reproduced as is)
_FLAGS = [
    ("voxel_shape", int, [128, 128, 128]), ("load", str, None), ("batch_size", int, 1), ("num_warmup_steps", int, 1024),
    ("save_freq", int, 1), ("eval_freq", int, 5), ("seed", int, 0), ("epochs", int, 200), ("num_descs", int, 4),
    ("saliency_vmin", float, None), ("lr", float, 1e-3), ("weight_decay", float, 0.00001), ("grad_max_norm", float, 2.0),
    ("xyz_pts_noise", float, 0.0), ("num_input_pts", int, 80000), ("num_output_pts", int, 400000), ("pointing_dim", int, 64),
    ("unet_f_maps", int, 16), ("unet_num_channels", int, 16), ("unet_num_groups", int, 8), ("unet_num_levels", int, 6),
    ("num_patches", int, 4), ("patch_mask_cutoff", float, 0.004), ("pts_feat_extractor_hidden_dim", int, 128),
    ("num_workers", int, 8), ("dr_pos", float, 0.1), ("dr_orn", float, 0.3), ("dr_scale", float, 0.1), ("device", str, "cuda"),
]  # fmt: skip
_SWITCHES = [("domain_randomization", True), ("use_pts_feat_extractor", True), ("subtract_mean_relevancy", True),
             ("offset_patch_mask", False), ("balance_positive_negative", False), ("balance_spatial_relations", True),
             ("always_replace_subsample_pts", False), ("balance_spatial_sampling", True), ("decoder_concat_xyz_pts", True),
             ("use_amp", False)]  # fmt: skip


def config_parser() -> ArgumentParser:
    p = ArgumentParser()
    p.add_argument("--file_path", type=str, required=True)
    for name, typ, default in _FLAGS:
        p.add_argument("--" + name, type=typ, default=default)
    for name, default in _SWITCHES:
        p.add_argument("--" + name, action="store_true", default=default)
    p.add_argument("--gpus", type=str, nargs="+", default="0")
    p.add_argument("--scene_bounds", type=list, default=[[-1.0, -1.0, -0.1], [1.0, 1.0, 1.9]])
    p.add_argument("--pointing_method", choices=["cosine_sim", "dot_product", "additive"], default="cosine_sim")
    p.add_argument("--saliency_config", choices=saliency_configs.keys(), default="ours")
    p.add_argument("--network_inputs", nargs="+", choices=["patch_masks", "saliency", "rgb", "tsdf"], default=["saliency"])
    p.add_argument("--lr_scheduler_type", default="cosine_with_restarts",
                   choices=["constant", "linear", "cosine", "cosine_with_restarts", "constant_with_warmup"])
    p.add_argument("--reduce_method", choices=["max", "mean"], default="max")
    return p


def is_main_process() -> bool:
    return dist.get_rank() == 0 if dist.is_initialized() else True


def seed_all(seed: int = 0) -> None:
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def get_n_params(model) -> int:
    return sum(int(p.numel()) for p in model.parameters())


def compute_grad_norm(net) -> float:
    """utils.compute_grad_norm (:321-327): L2 norm over every parameter gradient, as a python float."""
    sq = [p.grad.detach().double().pow(2).sum() for p in net.parameters() if p.grad is not None]
    return float(torch.stack(sq).sum().sqrt()) if sq else 0.0


# ---------------------------------------------------------------------------------------------------------------------
# checkpoints: {"net": state_dict, "optimizer": state_dict, "epochs": int}   (utils.py:533-541, 604-613)
# ---------------------------------------------------------------------------------------------------------------------
def strip_module_prefix(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The reference strips everything up to and including the first `module.` (utils.py:282-287); keys written without
    the DDP wrapper (no prefix) are accepted unchanged here (the reference's expression would map them to "")."""
    out = {}
    for k, v in state_dict.items():
        parts = k.split("module.")
        out["module.".join(parts[1:]) if len(parts) > 1 and parts[0] == "" else k] = v
    return out


def checkpoint_state(net, optimizer, epochs: int) -> dict:
    sd = net.state_dict()
    if dist.is_initialized():  # what DistributedDataParallel(net).state_dict() looks like (reference checkpoints)
        sd = {"module." + k: v for k, v in sd.items()}
    return {"net": sd, "optimizer": optimizer.state_dict(), "epochs": epochs}


def save_checkpoint(path: str, net, optimizer, epochs: int) -> None:
    torch.save(checkpoint_state(net, optimizer, epochs), path)


def get_net(load: Optional[str], lr: float, weight_decay: float, lr_scheduler_type: str, num_warmup_steps: int, epochs: int,
            seed: int, net_class, use_amp: bool, train_dataset=None, **kwargs):
    """-> (net, optimizer, lr_scheduler, start_epoch, scaler) like utils.get_net (:237-294)."""
    from transformers import get_scheduler

    seed_all(seed)
    device = kwargs["device"]
    batch_size = kwargs["batch_size"]
    kwargs["voxel_shape"] = tuple(kwargs["voxel_shape"])
    net = net_class(precise=not use_amp, **kwargs).to(device)
    logging.info(f"NUM PARAMS: {get_n_params(net)}")
    optimizer = Lamb(net.parameters(), lr=lr, betas=(0.9, 0.999), weight_decay=weight_decay, adam=False)
    lr_scheduler = get_scheduler(
        lr_scheduler_type, optimizer=optimizer, num_warmup_steps=num_warmup_steps,
        num_training_steps=epochs * (len(train_dataset) // batch_size) if train_dataset is not None else 1)
    start_epoch = 0
    if load is not None:
        logging.info(f"loading from {load}")
        ckpt = torch.load(load, map_location=device, weights_only=False)
        net.load_state_dict(strip_module_prefix(ckpt["net"]))
        optimizer.load_state_dict(ckpt["optimizer"])
        start_epoch = ckpt["epochs"]
    return net, optimizer, lr_scheduler, start_epoch, None


# ---------------------------------------------------------------------------------------------------------------------
def loop(net, loader, pbar, get_losses_fn: Callable, logger=None, optimizer=None, lr_scheduler=None, scaler=None,
         grad_max_norm: float = 1e5, device=torch.device("cuda"), **kwargs) -> pd.DataFrame:
    """One pass over `loader` (utils.loop :383-471): train branch when an optimiser is given (losses -> backward ->
    gradient all-reduce -> clip -> LAMB -> scheduler -> steps += 1 -> gradnorm), otherwise evaluation under no_grad;
    per-step stats are averaged over ranks, the per-item DataFrames gathered from all ranks."""
    epoch_stats: Dict[str, list] = {}
    frames = []
    for batch in loader:
        batch = {k: (v.to(device) if type(v) == torch.Tensor else v) for k, v in batch.items()}
        if optimizer:
            stats, detailed = get_losses_fn(net=net, batch=batch, **kwargs)
            optimizer.zero_grad(set_to_none=True)
            with _train.GradientBuckets() as buckets:
                stats["loss"].backward()
            _train.all_reduce_gradients(net.parameters(), skip=buckets.reduced)
            if isinstance(optimizer, Lamb):
                optimizer.step(max_grad_norm=grad_max_norm)
                total = optimizer.last_grad_norm
            else:
                total = _train.clip_grad_norm_(net.parameters(), grad_max_norm)
                optimizer.step()
                _train.bump_weights_epoch()
            if lr_scheduler is not None:
                lr_scheduler.step()
            net.steps += 1
            # what utils.compute_grad_norm reads after the reference's in-place clip
            stats["gradnorm"] = float(total * torch.clamp(grad_max_norm / (total + 1e-6), max=1.0))
        else:
            with torch.no_grad():
                stats, detailed = get_losses_fn(net=net, batch=batch, **kwargs)
        if dist.is_initialized():
            keys = sorted(stats.keys())
            vec = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=device)
            dist.all_reduce(vec)
            for k, v in zip(keys, (vec / dist.get_world_size()).tolist()):
                stats[k] = v
            gathered = [None] * dist.get_world_size()
            dist.all_gather_object(gathered, detailed)
            frames.extend(gathered)
        else:
            frames.append(detailed)
        for k, v in stats.items():
            v = float(v)
            epoch_stats.setdefault(k, []).append(v)
            if logger is not None and optimizer is not None:
                logger.add_scalar(k, v, int(net.steps))
        if pbar is not None:
            pbar.set_description("|".join(
                f" {k}: {float(v) * 100:.02f} " if any(s in k for s in ("iou", "precision", "recall")) else f" {k}: {float(v):.04e} "
                for k, v in stats.items()))
            pbar.update()
    if logger is not None and is_main_process():
        for k, v in epoch_stats.items():
            logger.add_scalar(f"{k}_mean", float(np.nanmean(v)), int(net.steps))
    frames = [f for f in frames if f is not None]
    return pd.concat(frames) if frames else pd.DataFrame()


def train(log: str, net, optimizer, lr_scheduler, training_detailed_stats, start_epoch: int, epochs: int, datasets: dict,
          loggers: dict, splits: dict, save_freq: int, eval_freq: int, num_workers: int, batch_size: int, get_losses_fn,
          use_amp: bool = False, **kwargs):
    """Epoch driver (utils.train :474-614): the train split every epoch, `unseen_instances` every `eval_freq` epochs,
    `latest.pth` after every split pass, `ckpt_<epoch>.pth` every `save_freq` epochs and at the end."""
    from torch.utils.data import DataLoader
    from torch.utils.data.distributed import DistributedSampler

    os.makedirs(log, exist_ok=True)
    if training_detailed_stats is None:
        training_detailed_stats = pd.DataFrame()
    for curr_epoch in range(start_epoch, epochs):
        if is_main_process():
            logging.info(f'{"=" * 10} EPOCH {curr_epoch} {"=" * 10}')
        for split, dataset in datasets.items():
            if split != "train" and (curr_epoch % eval_freq != 0 or split != "unseen_instances"):
                continue
            net.train(split == "train")
            sampler = None
            if dist.is_initialized():
                sampler = DistributedSampler(dataset=dataset, shuffle=split == "train", drop_last=split == "train")
                sampler.set_epoch(curr_epoch)
            loader = DataLoader(dataset=dataset, sampler=sampler, num_workers=num_workers,
                                shuffle=sampler is None and split == "train", batch_size=batch_size if split == "train" else 1,
                                persistent_workers=num_workers > 0)
            detailed = loop(net=net, loader=loader, get_losses_fn=get_losses_fn,
                            **{**kwargs, "logger": loggers.get(split), "optimizer": optimizer if split == "train" else None,
                               "lr_scheduler": lr_scheduler, "pbar": None, "detailed_analysis": False,
                               "cutoffs": [-1.0] if split == "train" else np.arange(-2.7, 0, 0.3)})
            if is_main_process():
                save_checkpoint(f"{log}/latest.pth", net, optimizer, curr_epoch + 1)
            detailed["epoch"] = [curr_epoch] * len(detailed)
            detailed["split"] = [split] * len(detailed)
            training_detailed_stats = pd.concat([training_detailed_stats, detailed])
            if is_main_process():
                training_detailed_stats.to_pickle(log + "/detailed_stats.pkl")
        if not is_main_process():
            continue
        if curr_epoch % save_freq != 0 and curr_epoch != epochs - 1:
            continue
        save_checkpoint(f"{log}/ckpt_{curr_epoch}.pth", net, optimizer, curr_epoch + 1)
        logging.info(f"Saved checkpoint to {log}/ckpt_{curr_epoch}.pth.")
    return training_detailed_stats


# ---------------------------------------------------------------------------------------------------------------------
# synthetic stand-ins for dataset.SceneCompletionDataset / ObjectLocalizationDataset (batch contract: SURVEY.md §8d)
# ---------------------------------------------------------------------------------------------------------------------
class _SyntheticBase(torch.utils.data.Dataset):
    def __init__(self, length=8, scene_bounds=((-1.0, -1.0, -0.1), (1.0, 1.0, 1.9)), num_input_pts=80000,
                 num_output_pts=400000, seed=0, **kwargs):
        self.length, self.n_in, self.n_out, self.seed = length, num_input_pts, num_output_pts, seed
        self.lo, self.hi = torch.tensor(scene_bounds[0]), torch.tensor(scene_bounds[1])

    def __len__(self):
        return self.length

    def _pts(self, g, *lead):
        return self.lo + (self.hi - self.lo) * torch.rand(*lead, 3, generator=g)


class SyntheticOVSSCDataset(_SyntheticBase):
    """dataset.py:975-1238 as consumed at train_ovssc.py:81-144."""

    def __init__(self, num_patches=4, **kwargs):
        super().__init__(**kwargs)
        self.P = num_patches

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 100003 + i)
        P = self.P
        return dict(input_xyz_pts=self._pts(g, self.n_in), input_feature_pts=torch.randn(P, self.n_in, 1, generator=g),
                    tsdf_vol=torch.ones(1), output_xyz_pts=self._pts(g, P, self.n_out),
                    output_label_pts=(torch.rand(P, self.n_out, generator=g) < 0.1).float(),
                    out_of_bounds_pts=torch.zeros(P, self.n_out, dtype=torch.bool),
                    out_of_frustum_pts_mask=torch.rand(P, self.n_out, generator=g) < 0.05,
                    semantic_class_features=torch.randn(P, 512, generator=g),
                    patch_labels=[f"object {k}" for k in range(P - 1)] + [""], scene_id=f"synthetic_{i}")


class SyntheticVOOLDataset(_SyntheticBase):
    """dataset.py:330-678 as consumed at train_vool.py:118-178."""

    RELATIONS = ["in", "behind", "in front of", "on the left of", "on the right of", "on"]

    def __init__(self, num_descs=4, **kwargs):
        super().__init__(**kwargs)
        self.D = num_descs

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 100003 + i)
        D = self.D
        sal = lambda: torch.randn(D, self.n_in, 1, generator=g)
        return dict(input_xyz_pts=self._pts(g, self.n_in), input_target_saliency_pts=sal(), input_reference_saliency_pts=sal(),
                    input_description_saliency_pts=sal(), tsdf_vol=torch.ones(1), output_xyz_pts=self._pts(g, D, self.n_out),
                    output_label_pts=(torch.rand(D, self.n_out, generator=g) < 0.1).float(),
                    out_of_bounds_pts=torch.zeros(D, self.n_out, dtype=torch.bool),
                    out_of_frustum_pts_mask=torch.zeros(D, self.n_out, dtype=torch.bool),
                    spatial_relation_name=[self.RELATIONS[(i + d) % 6] for d in range(D)],
                    target_obj_name=[f"target {d}" for d in range(D)], reference_obj_name=[f"reference {d}" for d in range(D)],
                    scene_id=f"synthetic_{i}")


def setup_experiment(args, net_class, dataset_class=None, **kwargs) -> dict:
    """utils.setup_experiment (:144-213) for this path: process group from the torchrun environment (one process per GPU,
    NCCL; the reference additionally sets NCCL_P2P_DISABLE=1, utils.py:132 — not here), network / optimiser / scheduler
    through `get_net`, datasets from `dataset_class` (a stand-in for dataset.py's HDF5 loaders: `--file_path synthetic`
    selects the seeded synthetic batches)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if args.device == "cuda":
            torch.cuda.set_device(local)
            args.device = f"cuda:{local}"
        dist.init_process_group(backend="nccl" if args.device.startswith("cuda") else "gloo", init_method="env://")
    logging.getLogger().setLevel(logging.INFO if is_main_process() else logging.ERROR)
    if dataset_class is None or args.file_path != "synthetic":
        if dataset_class is None:
            raise NotImplementedError("the HDF5 scene datasets (dataset.py) are outside this package's hot path: pass the "
                                      "reference's dataset class as `dataset_class`, or use --file_path synthetic")
    ds_kwargs = {**vars(args), **kwargs}
    datasets = {"train": dataset_class(**ds_kwargs)}
    net, optimizer, lr_scheduler, start_epoch, scaler = get_net(train_dataset=datasets["train"], net_class=net_class, **vars(args))
    stats = None
    if os.path.exists(args.log + "/detailed_stats.pkl"):
        stats = pd.read_pickle(args.log + "/detailed_stats.pkl")
    return {"splits": {"train": []}, "loggers": {}, "datasets": datasets, "net": net, "scaler": scaler, "optimizer": optimizer,
            "lr_scheduler": lr_scheduler, "start_epoch": start_epoch, "training_detailed_stats": stats}
