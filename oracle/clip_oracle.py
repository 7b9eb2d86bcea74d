"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain torch fp32 on CPU + autograd) of the reference's relevancy path.

Nothing here is shipped or timed as the product; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs import it.  Pinning: `oracle/gen_golden.py` runs the UNMODIFIED reference (imported from
/root/reference through oracle/ref_import.py) on the same seeded weights/inputs and asserts this restatement agrees
with it; the reference outputs are committed under tests/golden/ and re-checked by the CPU test-suite
(tests/test_oracle_golden.py).  The reference itself has no numeric tests/KATs for this path.

Each function cites the reference lines it follows.  State dicts use the reference's key names.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F


def convert_weights_values(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """build_model -> convert_weights -> load_state_dict -> .float(): Linear/Conv/MHA weights+biases and the two
    projection matrices pass through fp16 (CLIP/clip/model_explainability.py:501-527, 600-602;
    clip_explainability.py:163-168)."""
    suffixes = ("conv1.weight", "attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight",
                "attn.out_proj.bias", "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias")  # fmt: skip
    out = {}
    for k, v in sd.items():
        if k.startswith("__"):
            continue
        v = v.detach().float()
        if k.endswith(suffixes) or k in ("visual.proj", "text_projection"):
            v = v.half().float()
        out[k] = v
    return out


def _ln(x, w, b):
    # LayerNorm subclass computes in fp32 (model_explainability.py:188-194)
    return F.layer_norm(x.float(), (x.shape[-1],), w, b, 1e-5)


def _positional_quirk(pos: torch.Tensor, target_len: int) -> torch.Tensor:
    # interpolate_positional_emb (CLIP/clip/auxiliary.py:24-38): index i reads row i/(T/50)
    out = torch.zeros(target_len, pos.shape[1])
    for i in range(target_len):
        i3 = float(i) / (target_len / 50)
        i1, i2 = math.floor(i3), math.ceil(i3)
        out[i] = torch.lerp(pos[i1], pos[i2], i3 - i1) if i2 < len(pos) else pos[-1]
    return out


def _block(sd, prefix, x, heads, mask=None, probs_out: List[torch.Tensor] | None = None):
    """ResidualAttentionBlock.forward (model_explainability.py:252-255) with multi_head_attention_forward
    (auxiliary.py:129 in-proj, :207 q scaling, :260-264 head split, :307 QK^T, :312-316 mask, :326 softmax,
    :334 hook, :337 A·V, :339-340 out-proj). x is [N, T, d]."""
    N, T, d = x.shape
    hd = d // heads
    h = _ln(x, sd[prefix + "ln_1.weight"], sd[prefix + "ln_1.bias"])
    qkv = F.linear(h, sd[prefix + "attn.in_proj_weight"], sd[prefix + "attn.in_proj_bias"])
    q, k, v = qkv.chunk(3, dim=-1)
    q = q * (float(hd) ** -0.5)
    q, k, v = (t.reshape(N, T, heads, hd).permute(0, 2, 1, 3) for t in (q, k, v))
    s = q @ k.transpose(-1, -2)
    if mask is not None:
        s = s + mask
    a = F.softmax(s, dim=-1)
    if probs_out is not None:
        probs_out.append(a)
    o = (a @ v).permute(0, 2, 1, 3).reshape(N, T, d)
    x = x + F.linear(o, sd[prefix + "attn.out_proj.weight"], sd[prefix + "attn.out_proj.bias"])
    h = _ln(x, sd[prefix + "ln_2.weight"], sd[prefix + "ln_2.bias"])
    u = F.linear(h, sd[prefix + "mlp.c_fc.weight"], sd[prefix + "mlp.c_fc.bias"])
    u = u * torch.sigmoid(1.702 * u)  # QuickGELU (:197-199)
    return x + F.linear(u, sd[prefix + "mlp.c_proj.weight"], sd[prefix + "mlp.c_proj.bias"])


def encode_image(sd, tiles: torch.Tensor, probs_out: List[torch.Tensor] | None = None) -> torch.Tensor:
    """VisionTransformer.forward (model_explainability.py:324-355). tiles [B,3,R,R] fp32 -> [B,E]."""
    w = sd["visual.conv1.weight"]
    width, patch = w.shape[0], w.shape[-1]
    heads = width // 64
    x = F.conv2d(tiles, w, stride=patch)
    x = x.reshape(x.shape[0], width, -1).permute(0, 2, 1)
    cls = sd["visual.class_embedding"] + torch.zeros(x.shape[0], 1, width)
    x = torch.cat([cls, x], dim=1)
    pos = sd["visual.positional_embedding"]
    if x.shape[1] != 50:
        pos = _positional_quirk(pos, x.shape[1])
    x = x + pos
    x = _ln(x, sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"])
    layers = len([k for k in sd if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
    for i in range(layers):
        x = _block(sd, f"visual.transformer.resblocks.{i}.", x, heads, probs_out=probs_out)
    x = _ln(x[:, 0, :], sd["visual.ln_post.weight"], sd["visual.ln_post.bias"])
    return x @ sd["visual.proj"]


def encode_text(sd, tokens: torch.Tensor) -> torch.Tensor:
    """CLIP.encode_text (model_explainability.py:468-482) with the causal mask of :452-458."""
    x = sd["token_embedding.weight"][tokens] + sd["positional_embedding"]
    width = x.shape[-1]
    ctx = x.shape[1]
    mask = torch.full((ctx, ctx), float("-inf")).triu_(1)
    layers = len(set(k.split(".")[2] for k in sd if k.startswith("transformer.resblocks")))
    for i in range(layers):
        x = _block(sd, f"transformer.resblocks.{i}.", x, width // 64, mask=mask)
    x = _ln(x, sd["ln_final.weight"], sd["ln_final.bias"])
    return x[torch.arange(x.shape[0]), tokens.argmax(dim=-1)] @ sd["text_projection"]


def zeroshot_weights(sd, tokens: torch.Tensor, n_classes: int, n_templates: int) -> torch.Tensor:
    """zeroshot_classifier (CLIP/clip/clip_gradcam.py:12-27): per-template L2 norm, mean over templates -> [E,P]."""
    with torch.no_grad():
        e = encode_text(sd, tokens).view(n_classes, n_templates, -1)
        e = e / e.norm(dim=-1, keepdim=True)
        return e.mean(dim=1).T.contiguous()


def relevancy(sd, tiles: torch.Tensor, W: torch.Tensor, num_layers: int = 10, positive_attn_only: bool = True,
              return_logits: bool = False):
    """ClipGradcam.forward + interpret (clip_gradcam.py:58-132): W [E,P] -> [P,B,g,g].
    Gradients of sum_b logit[b,l] w.r.t. every kept attention tensor are taken with one autograd.grad call per
    label (the reference issues one call per (label, block); same graph, same values)."""
    probs: List[torch.Tensor] = []
    with torch.enable_grad():
        # the weights in `sd` are plain tensors: make the graph differentiable through the input instead
        f = encode_image(sd, tiles.detach().clone().requires_grad_(True), probs_out=probs)
        f = f / f.norm(dim=-1, keepdim=True)
        logits = 100.0 * f @ W
        B, P = logits.shape
        T = probs[0].shape[-1]
        kept = [i for i in range(len(probs)) if i > num_layers]
        grads = [torch.autograd.grad(logits[:, l].sum(), [probs[i] for i in kept], retain_graph=True) for l in range(P)]
    R = torch.eye(T).repeat(P, B, 1, 1)
    for bi, i in enumerate(kept):
        grad = torch.stack([grads[l][bi].detach() for l in range(P)])  # [P,B,H,T,T]
        cam = grad * probs[i].detach()[None]
        if positive_attn_only:
            cam = cam.clamp(min=0)
        cam = cam.mean(dim=2)  # heads
        R = R + torch.bmm(cam.reshape(P * B, T, T), R.reshape(P * B, T, T)).view(P, B, T, T)
    g = int(np.sqrt(T - 1))
    rel = R[:, :, 0, 1:].reshape(P, B, g, g)
    if return_logits:
        return rel, logits.detach()
    return rel


# ---------------------------------------------------------------------------------------------------------
# tiling pyramid and assembly
# ---------------------------------------------------------------------------------------------------------
def preprocess_tile(tile_u8: np.ndarray, n_px: int = 224) -> torch.Tensor:
    """_transform (clip_explainability.py:98-108): Resize(224, bicubic) -> CenterCrop(n_px) -> ToTensor -> Normalize."""
    from PIL import Image

    img = Image.fromarray(tile_u8).convert("RGB")
    w, h = img.size
    if w <= h:
        nw, nh = 224, int(224 * h / w)
    else:
        nw, nh = int(224 * w / h), 224
    img = img.resize((nw, nh), Image.BICUBIC)
    arr = torch.from_numpy(np.array(img, dtype=np.uint8)).permute(2, 0, 1).float().div(255)
    if (nh, nw) != (n_px, n_px):
        # CenterCrop pads with zeros when the image is smaller than the crop (torchvision semantics)
        if nh < n_px or nw < n_px:
            pl, pt = max((n_px - nw) // 2, 0), max((n_px - nh) // 2, 0)
            pr, pb = max((n_px - nw + 1) // 2, 0), max((n_px - nh + 1) // 2, 0)
            arr = F.pad(arr, (pl, pr, pt, pb))
            nh, nw = arr.shape[1:]
        top, left = int(round((nh - n_px) / 2.0)), int(round((nw - n_px) / 2.0))
        arr = arr[:, top : top + n_px, left : left + n_px]
    mean = torch.tensor((0.48145466, 0.4578275, 0.40821073)).view(3, 1, 1)
    std = torch.tensor((0.26862954, 0.26130258, 0.27577711)).view(3, 1, 1)
    return (arr - mean) / std


def enumerate_tiles(shape, cropping_augmentations: Sequence[dict], n_images: int = 1):
    """Tile enumeration order of ClipWrapper.create_tiles (CLIP/clip/__init__.py:257-273): images -> crop sizes ->
    y (column offset, outer) -> x (row offset, inner). Returns int array [n,3] = (row0, col0, size)."""
    H, W = shape[:2]
    out = []
    for _ in range(n_images):
        for aug in cropping_augmentations:
            ts, st = aug["tile_size"], aug["stride"]
            for y in np.arange(0, W - ts + 1, st):
                if y >= H:
                    continue
                for x in np.arange(0, H - ts + 1, st):
                    if x >= W:
                        continue
                    out.append((int(x), int(y), int(ts)))
    return np.array(out, dtype=np.int32).reshape(-1, 3)


def assemble(rel: torch.Tensor, tile_desc: np.ndarray, size_order: Sequence[int], H: int, W: int) -> torch.Tensor:
    """Up-sample + overlap-add + normalise (CLIP/clip/__init__.py:205-236): fp16 accumulators, tiles added in
    creation order per tile size (sizes visited ascending like np.unique, :207), counts start at 1e-5 (:249-253),
    final sum over sizes in cropping_augmentations order (:230-233). rel is [P, n, g, g] fp32."""
    P = rel.shape[0]
    outputs = {s: torch.zeros(P, H, W).half() for s in size_order}
    counts = {s: torch.zeros(H, W) + 1e-5 for s in size_order}
    for r0, c0, s in tile_desc:
        counts[int(s)][r0 : r0 + s, c0 : c0 + s] += 1
    for s in np.unique(tile_desc[:, 2]):
        idx = np.nonzero(tile_desc[:, 2] == s)[0]
        for start in range(0, len(idx), 32):
            chunk = idx[start : start + 32]
            up = F.interpolate(rel[:, chunk], size=int(s), mode="bilinear", align_corners=False)
            for k, ti in enumerate(chunk):
                r0, c0, _ = tile_desc[ti]
                outputs[int(s)][:, r0 : r0 + s, c0 : c0 + s] += up[:, k]
    return sum(outputs[s].float() / counts[s] for s in size_order) / len(size_order)


def get_clip_saliency(sd, img: np.ndarray, W: torch.Tensor, cropping_augmentations, positive_attn_only=True,
                      horizontal_flipping=False, jittered_images: Sequence[np.ndarray] = (), tile_batch=32,
                      num_layers: int = 10, n_px: int = 224):
    """get_clip_saliency_convolve (CLIP/clip/__init__.py:135-236) for explicit zero-shot weights W [E,P].
    `jittered_images` are the ColorJitter copies the reference would draw (:246-247) — supplied by the caller so
    that both sides see the same random augmentation."""
    images = [img] + list(jittered_images)
    desc = enumerate_tiles(img.shape, cropping_augmentations, n_images=len(images))
    per_img = len(desc) // len(images)
    tiles = torch.stack([
        preprocess_tile(images[i // per_img][r0 : r0 + s, c0 : c0 + s], n_px) for i, (r0, c0, s) in enumerate(desc)
    ])  # fmt: skip

    def run(t):
        return torch.cat([relevancy(sd, t[i : i + tile_batch], W, num_layers, positive_attn_only)
                          for i in range(0, len(t), tile_batch)], dim=1)  # fmt: skip

    rel = run(tiles)
    if horizontal_flipping:
        rel = (rel + run(tiles.flip(-1)).flip(-1)) / 2
    size_order = [a["tile_size"] for a in cropping_augmentations]
    return assemble(rel, desc, size_order, img.shape[0], img.shape[1])
