"""CLIP geometry, state-dict handling and the device-side weight pack consumed by the CUDA engine.

State dicts use the reference's key names (`visual.transformer.resblocks.3.attn.in_proj_weight`, ...;
reference: CLIP/clip/model_explainability.py:360-482, build_model :530-602) so an OpenAI checkpoint
(`~/.cache/clip/ViT-*.pt`) or a seeded synthetic one is loaded the same way.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

# (embed_dim, image_resolution, vision_layers, vision_width, vision_patch, context, vocab, text_width, text_heads, text_layers)
CLIP_GEOMETRY = {
    "ViT-B/32": (512, 224, 12, 768, 32, 77, 49408, 512, 8, 12),
    "ViT-B/16": (512, 224, 12, 768, 16, 77, 49408, 512, 8, 12),
    "ViT-L/14": (768, 224, 24, 1024, 14, 77, 49408, 768, 12, 12),
    "ViT-L/14@336px": (768, 336, 24, 1024, 14, 77, 49408, 768, 12, 12),
}


def available_models() -> List[str]:
    return list(CLIP_GEOMETRY.keys())


def synthetic_clip_state_dict(name: str, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded random-init CLIP weights with the reference's parameter names/shapes (no checkpoints are reachable
    offline). Standard deviations follow CLIP.initialize_parameters (model_explainability.py:419-450) for the text
    tower and torch's default Linear / MultiheadAttention scales for the vision tower; biases are small non-zero
    values so that every bias path is exercised."""
    E, res, vl, vw, vp, ctx, vocab, tw, th, tl = CLIP_GEOMETRY[name]
    g = torch.Generator().manual_seed(seed)

    def nrm(*shape, std):
        return torch.randn(*shape, generator=g) * std

    def uni(*shape, bound):
        return (torch.rand(*shape, generator=g) * 2 - 1) * bound

    sd: Dict[str, torch.Tensor] = {}
    scale = vw**-0.5
    sd["visual.class_embedding"] = nrm(vw, std=scale)
    sd["visual.positional_embedding"] = nrm((res // vp) ** 2 + 1, vw, std=scale)
    sd["visual.proj"] = nrm(vw, E, std=scale)
    sd["visual.conv1.weight"] = uni(vw, 3, vp, vp, bound=(3 * vp * vp) ** -0.5)
    for nm in ("ln_pre", "ln_post"):
        sd[f"visual.{nm}.weight"] = 1.0 + nrm(vw, std=0.02)
        sd[f"visual.{nm}.bias"] = nrm(vw, std=0.02)

    def tower(prefix, width, layers, attn_std, proj_std, fc_std):
        for i in range(layers):
            p = f"{prefix}transformer.resblocks.{i}."
            sd[p + "attn.in_proj_weight"] = nrm(3 * width, width, std=attn_std)
            sd[p + "attn.in_proj_bias"] = nrm(3 * width, std=0.01)
            sd[p + "attn.out_proj.weight"] = nrm(width, width, std=proj_std)
            sd[p + "attn.out_proj.bias"] = nrm(width, std=0.01)
            sd[p + "ln_1.weight"] = 1.0 + nrm(width, std=0.02)
            sd[p + "ln_1.bias"] = nrm(width, std=0.02)
            sd[p + "mlp.c_fc.weight"] = nrm(4 * width, width, std=fc_std)
            sd[p + "mlp.c_fc.bias"] = nrm(4 * width, std=0.01)
            sd[p + "mlp.c_proj.weight"] = nrm(width, 4 * width, std=proj_std)
            sd[p + "mlp.c_proj.bias"] = nrm(width, std=0.01)
            sd[p + "ln_2.weight"] = 1.0 + nrm(width, std=0.02)
            sd[p + "ln_2.bias"] = nrm(width, std=0.02)

    # vision tower: torch defaults are ~U(-1/sqrt(in), 1/sqrt(in)) => std = 1/sqrt(3 in)
    tower("visual.", vw, vl, attn_std=(2.0 / (4 * vw)) ** 0.5, proj_std=(3 * vw) ** -0.5, fc_std=(3 * vw) ** -0.5)
    # text tower: CLIP.initialize_parameters
    tower("", tw, tl, attn_std=tw**-0.5, proj_std=(tw**-0.5) * ((2 * tl) ** -0.5), fc_std=(2 * tw) ** -0.5)
    sd["token_embedding.weight"] = nrm(vocab, tw, std=0.02)
    sd["positional_embedding"] = nrm(ctx, tw, std=0.01)
    sd["ln_final.weight"] = 1.0 + nrm(tw, std=0.02)
    sd["ln_final.bias"] = nrm(tw, std=0.02)
    sd["text_projection"] = nrm(tw, E, std=tw**-0.5)
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07))
    return sd


_FP16_SUFFIXES = (
    "conv1.weight",
    "attn.in_proj_weight",
    "attn.in_proj_bias",
    "attn.out_proj.weight",
    "attn.out_proj.bias",
    "mlp.c_fc.weight",
    "mlp.c_fc.bias",
    "mlp.c_proj.weight",
    "mlp.c_proj.bias",
)


def apply_convert_weights(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """What build_model -> convert_weights -> load_state_dict -> .float() does to the values (reference:
    model_explainability.py:501-527,600-602; clip_explainability.py:163-168): Conv/Linear/MHA weights and biases
    plus `proj` / `text_projection` are rounded to fp16; LayerNorm, embeddings and positional tables stay fp32."""
    out = {}
    for k, v in sd.items():
        v = v.detach().float()
        if k.endswith(_FP16_SUFFIXES) or k in ("visual.proj", "text_projection"):
            v = v.half().float()
        out[k] = v
    return out


def interpolate_positional_embedding(pos: torch.Tensor, target_len: int) -> torch.Tensor:
    """The reference's positional "interpolation" for T != 50 (CLIP/clip/auxiliary.py:24-38): token i reads the
    table at i / (T / 50) — i.e. only the first ~50 rows are ever used — lerping between floor and ceil rows.
    Reproduced as is (load-bearing quirk)."""
    n = pos.shape[0]
    out = torch.zeros(target_len, pos.shape[1], dtype=pos.dtype)
    for i in range(target_len):
        i3 = float(i) / (target_len / 50)
        i1, i2 = math.floor(i3), math.ceil(i3)
        if i2 < n:
            out[i] = torch.lerp(pos[i1], pos[i2], i3 - i1)
        else:
            out[i] = pos[-1]
    return out


@dataclass
class BlockWeights:
    ln1_g: torch.Tensor
    ln1_b: torch.Tensor
    w_in: torch.Tensor  # [3d, d] fp16
    b_in: torch.Tensor
    w_out: torch.Tensor  # [d, d]
    b_out: torch.Tensor
    ln2_g: torch.Tensor
    ln2_b: torch.Tensor
    w_fc: torch.Tensor  # [4d, d]
    b_fc: torch.Tensor
    w_proj: torch.Tensor  # [d, 4d]
    b_proj: torch.Tensor
    # transposed copies = B operands of the dgrad GEMMs (only for blocks that take part in the rollout)
    w_inT: Optional[torch.Tensor] = None  # [d, 3d]
    w_outT: Optional[torch.Tensor] = None  # [d, d]
    w_fcT: Optional[torch.Tensor] = None  # [d, 4d]
    w_projT: Optional[torch.Tensor] = None  # [4d, d]


@dataclass
class TowerWeights:
    width: int
    heads: int
    blocks: List[BlockWeights] = field(default_factory=list)


@dataclass
class ClipDeviceWeights:
    name: str
    embed_dim: int
    input_resolution: int
    patch: int
    grid: int
    tokens: int
    kpad: int
    context_length: int
    visual: TowerWeights
    text: TowerWeights
    conv_w: torch.Tensor  # [d, kpad] fp16
    cls: torch.Tensor
    pos: torch.Tensor  # [T, d] table actually added (quirk applied)
    ln_pre_g: torch.Tensor
    ln_pre_b: torch.Tensor
    ln_post_g: torch.Tensor
    ln_post_b: torch.Tensor
    proj: torch.Tensor  # [d, E] fp16 (B operand of dy = df proj^T)
    projT: torch.Tensor  # [E, d] fp16 (B operand of f = y proj)
    token_embedding: torch.Tensor
    text_pos: torch.Tensor
    ln_final_g: torch.Tensor
    ln_final_b: torch.Tensor
    text_projT: torch.Tensor  # [E, w] fp16


def _tower(sd, prefix, width, layers, device, first_bwd_block) -> TowerWeights:
    tw = TowerWeights(width=width, heads=width // 64)
    f16 = lambda t: t.to(device=device, dtype=torch.float16).contiguous()
    f32 = lambda t: t.to(device=device, dtype=torch.float32).contiguous()
    for i in range(layers):
        p = f"{prefix}transformer.resblocks.{i}."
        b = BlockWeights(
            ln1_g=f32(sd[p + "ln_1.weight"]), ln1_b=f32(sd[p + "ln_1.bias"]),
            w_in=f16(sd[p + "attn.in_proj_weight"]), b_in=f32(sd[p + "attn.in_proj_bias"]),
            w_out=f16(sd[p + "attn.out_proj.weight"]), b_out=f32(sd[p + "attn.out_proj.bias"]),
            ln2_g=f32(sd[p + "ln_2.weight"]), ln2_b=f32(sd[p + "ln_2.bias"]),
            w_fc=f16(sd[p + "mlp.c_fc.weight"]), b_fc=f32(sd[p + "mlp.c_fc.bias"]),
            w_proj=f16(sd[p + "mlp.c_proj.weight"]), b_proj=f32(sd[p + "mlp.c_proj.bias"]),
        )  # fmt: skip
        if first_bwd_block is not None and i >= first_bwd_block:
            b.w_inT = f16(sd[p + "attn.in_proj_weight"].t())
            b.w_outT = f16(sd[p + "attn.out_proj.weight"].t())
            b.w_fcT = f16(sd[p + "mlp.c_fc.weight"].t())
            b.w_projT = f16(sd[p + "mlp.c_proj.weight"].t())
        tw.blocks.append(b)
    return tw


def pack_clip_weights(name: str, state_dict: Dict[str, torch.Tensor], device, first_rollout_block: int = 11) -> ClipDeviceWeights:
    """Host -> device weight pack (one-off; mirrors `load`/`build_model`, reference clip_explainability.py:116-169)."""
    sd = apply_convert_weights(state_dict)
    vw = sd["visual.conv1.weight"].shape[0]
    patch = sd["visual.conv1.weight"].shape[-1]
    vl = len([k for k in sd if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
    grid_ckpt = round((sd["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    res = patch * grid_ckpt
    E = sd["text_projection"].shape[1]
    ctx = sd["positional_embedding"].shape[0]
    tw = sd["ln_final.weight"].shape[0]
    tl = len(set(k.split(".")[2] for k in sd if k.startswith("transformer.resblocks")))
    # the reference preprocesses every tile to 224 px regardless of the checkpoint (clip_explainability.py:106-107)
    # and CenterCrops to the model resolution; tokens follow from the model resolution
    g = res // patch
    T = g * g + 1
    kp = 3 * patch * patch
    kpad = (kp + 63) // 64 * 64
    conv = torch.zeros(vw, kpad)
    conv[:, :kp] = sd["visual.conv1.weight"].reshape(vw, kp)
    pos = sd["visual.positional_embedding"]
    if T != 50:
        pos = interpolate_positional_embedding(pos, T)
    f16 = lambda t: t.to(device=device, dtype=torch.float16).contiguous()
    f32 = lambda t: t.to(device=device, dtype=torch.float32).contiguous()
    return ClipDeviceWeights(
        name=name, embed_dim=E, input_resolution=res, patch=patch, grid=g, tokens=T, kpad=kpad, context_length=ctx,
        visual=_tower(sd, "visual.", vw, vl, device, first_rollout_block),
        text=_tower(sd, "", tw, tl, device, None),
        conv_w=f16(conv), cls=f32(sd["visual.class_embedding"]), pos=f32(pos),
        ln_pre_g=f32(sd["visual.ln_pre.weight"]), ln_pre_b=f32(sd["visual.ln_pre.bias"]),
        ln_post_g=f32(sd["visual.ln_post.weight"]), ln_post_b=f32(sd["visual.ln_post.bias"]),
        proj=f16(sd["visual.proj"]), projT=f16(sd["visual.proj"].t()),
        token_embedding=f32(sd["token_embedding.weight"]), text_pos=f32(sd["positional_embedding"]),
        ln_final_g=f32(sd["ln_final.weight"]), ln_final_b=f32(sd["ln_final.bias"]),
        text_projT=f16(sd["text_projection"].t()),
    )  # fmt: skip


def load_state_dict(name: str, download_root: Optional[str] = None, seed: Optional[int] = None,
                    allow_synthetic: Optional[bool] = None) -> Dict[str, torch.Tensor]:
    """Checkpoint lookup mirroring `load` (reference clip_explainability.py:116-169) minus the network download:
    a path, or `~/.cache/clip/<file>.pt`.  When neither exists the reference would download or raise; here (no network)
    this RAISES too unless the caller opted into seeded synthetic weights of the same geometry — explicitly, by passing
    `seed=` / `allow_synthetic=True` or setting SEMABS_B200_ALLOW_SYNTHETIC=1 (tests and bench.py do; a user asking for
    relevancy maps must never silently get maps of a random model)."""
    fname = {"ViT-B/32": "ViT-B-32.pt", "ViT-B/16": "ViT-B-16.pt", "ViT-L/14": "ViT-L-14.pt",
             "ViT-L/14@336px": "ViT-L-14-336px.pt"}  # fmt: skip
    path = None
    if os.path.isfile(name):
        path = name
    elif name in fname:
        cand = os.path.join(download_root or os.path.expanduser("~/.cache/clip"), fname[name])
        if os.path.isfile(cand):
            path = cand
    else:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    if path is None:
        if allow_synthetic is None:
            allow_synthetic = seed is not None or os.environ.get("SEMABS_B200_ALLOW_SYNTHETIC", "0") == "1"
        if not allow_synthetic:
            raise RuntimeError(
                f"CLIP checkpoint for {name} not found under {download_root or os.path.expanduser('~/.cache/clip')} "
                "(no network download here). Place the checkpoint there, or opt into seeded random-init weights of the "
                "same geometry with seed=<int> / allow_synthetic=True / SEMABS_B200_ALLOW_SYNTHETIC=1.")
        import warnings

        warnings.warn(f"semabs_b200: using SYNTHETIC seeded random-init weights for {name} (no checkpoint found); "
                      "relevancy maps are meaningless except for parity / throughput measurements", stacklevel=2)
        sd = synthetic_clip_state_dict(name, 0 if seed is None else seed)
        sd["__synthetic__"] = torch.tensor(1)
        return sd
    try:
        sd = torch.jit.load(path, map_location="cpu").state_dict()
    except RuntimeError:
        sd = torch.load(path, map_location="cpu", weights_only=True)
        if "state_dict" in sd:
            sd = sd["state_dict"]
    return {k: v for k, v in sd.items() if k not in ("input_resolution", "context_length", "vocab_size")}
