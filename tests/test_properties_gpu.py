"""Size-independent properties at BASELINE.json's FULL sizes (where the CPU oracle would take hours), plus the batching
edge cases of the public API (ragged tile batches, label chunks above prompt_batch_size, distractor labels).

Relevancy (configs[1]: ViT-L/14, 336^2, 285-tile pyramid, 16 labels): a map must not depend on how tiles and labels are
batched, permuting the labels permutes the maps, positive-only relevance is non-negative, repeated runs are identical.
Voxel UNet (configs[2]: 128^3 x 32 ch, batch 4): a grid's output must not depend on its batch neighbours; repeated runs
are identical."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
PYRAMID = [{"tile_size": s, "stride": s // 4} for s in (336, 224, 168, 112, 84)]
LABELS16 = ["basketball jersey", "nintendo switch", "television", "ping pong table", "vase", "fireplace",
            "abstract painting of a vespa", "carpet", "wall", "microwave", "cabinet", "fire extinguisher", "mirror",
            "woven chair", "globe", "leather sofa"]  # fmt: skip
PROMPT = "a photograph of a {} in a home."


def _maxrel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


def test_relevancy_full_config_batching_and_label_permutation():
    from semabs_b200.clip import ClipWrapper

    ClipWrapper.reset()
    ClipWrapper("ViT-L/14", "cuda", seed=0)
    img = np.random.default_rng(0).integers(0, 256, (336, 336, 3), dtype=np.uint8)
    cfg = dict(distractor_labels={}, horizontal_flipping=False, augmentations=0, positive_attn_only=True, cropping_augmentations=PYRAMID)
    a, _ = ClipWrapper.get_clip_saliency(img=img, text_labels=np.array(LABELS16), prompts=[PROMPT], tile_batch_size=95, **cfg)
    assert a.shape == (16, 336, 336) and torch.isfinite(a).all() and (a >= 0).all()
    a2, _ = ClipWrapper.get_clip_saliency(img=img, text_labels=np.array(LABELS16), prompts=[PROMPT], tile_batch_size=95, **cfg)
    assert _maxrel(a2, a) < 1e-6  # run-to-run stable
    # ragged tile batches (285 = 4 x 64 + 29) and label chunks of 5 (16 = 3 x 5 + 1)
    b, _ = ClipWrapper.get_clip_saliency(img=img, text_labels=np.array(LABELS16), prompts=[PROMPT], tile_batch_size=64,
                                         prompt_batch_size=5, **cfg)
    # the per-tile relevance agrees to ~1e-7; the assembly then accumulates in fp16 like the reference (__init__.py:149-153),
    # where such a difference can flip one rounding: tolerance = one fp16 ulp of the accumulator (4.9e-4)
    assert _maxrel(b, a) < 5e-4, _maxrel(b, a)
    # label permutation -> permuted maps
    perm = np.random.default_rng(1).permutation(16)
    c, _ = ClipWrapper.get_clip_saliency(img=img, text_labels=np.array(LABELS16)[perm], prompts=[PROMPT], tile_batch_size=95, **cfg)
    assert _maxrel(c, a[perm]) < 5e-4
    ClipWrapper.reset()


def test_more_labels_than_prompt_batch_and_distractors():
    """33 labels with the reference's default prompt_batch_size = 32 (two engine passes), and the distractor branch
    of get_clip_saliency (__init__.py:125-131: maps -= mean of the distractor maps), on ViT-B/32."""
    from semabs_b200.clip import ClipWrapper

    ClipWrapper.reset()
    ClipWrapper("ViT-B/32", "cuda", seed=0)
    img = np.random.default_rng(2).integers(0, 256, (96, 96, 3), dtype=np.uint8)
    cfg = dict(horizontal_flipping=True, augmentations=0, positive_attn_only=True,
               cropping_augmentations=[{"tile_size": 96, "stride": 24}, {"tile_size": 48, "stride": 12}])
    labels = [f"object number {i}" for i in range(33)]
    full, feats = ClipWrapper.get_clip_saliency(img=img, text_labels=np.array(labels), prompts=[PROMPT], distractor_labels={}, **cfg)
    assert full.shape == (33, 96, 96) and feats.shape[0] == 33
    part, _ = ClipWrapper.get_clip_saliency(img=img, text_labels=np.array(labels[30:]), prompts=[PROMPT], distractor_labels={}, **cfg)
    assert _maxrel(full[30:], part) < 5e-4
    d = {"object number 1", "object number 2", "object number 31"}  # the last one is also a target: removed from the set
    with_d, _ = ClipWrapper.get_clip_saliency(img=img, text_labels=np.array(labels[30:]), prompts=[PROMPT], distractor_labels=d, **cfg)
    expect = part - full[1:3].mean(dim=0)
    assert _maxrel(with_d, expect) < 1e-3
    ClipWrapper.reset()


@torch.no_grad()
def test_unet_full_size_batch_independence():
    from semabs_b200.unet3d import ResidualUNet3D

    torch.manual_seed(0)
    m = ResidualUNet3D(in_channels=32, out_channels=32, f_maps=32, num_groups=8, num_levels=6).cuda()
    x = torch.randn(4, 32, 128, 128, 128, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    y = m(x).clone()
    assert torch.isfinite(y).all()
    # run-to-run stable: the only order-dependent arithmetic is the fp64 atomics of the GroupNorm statistics; with the GroupNorm
    # folded into the level-0 convolutions their last-bit differences reach the weights (fp16 hi | lo images of W gamma rstd):
    # measured 1.1e-6
    assert _maxrel(m(x), y) < 5e-6
    y2 = m(x[2:3].contiguous())
    err = _maxrel(y2, y[2:3])
    assert err < 1e-5, err  # GroupNorm is per sample: a grid must not see its batch neighbours
