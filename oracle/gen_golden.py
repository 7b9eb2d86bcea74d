"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden [clip|config0|unet|net|all]
For every fixture it also asserts that the oracle restatement (oracle/*_oracle.py) reproduces the reference
output, which is what pins the oracle.  Inputs and weights are regenerated from seeds by the tests, so only the
reference OUTPUTS (plus checksums of the seeded inputs) are stored.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle import clip_oracle, ref_import  # noqa: E402

LABELS4 = ["basketball jersey", "nintendo switch", "television", "ping pong table"]
PROMPT = "a photograph of a {} in a home."


def synth_image(seed, h, w):
    """Smooth-ish random RGB image (blocks + noise) so crops at different scales differ."""
    rng = np.random.default_rng(seed)
    coarse = rng.integers(0, 256, (max(h // 16, 1), max(w // 16, 1), 3)).astype(np.float32)
    img = np.kron(coarse, np.ones((16, 16, 1), np.float32))[:h, :w]
    img = 0.7 * img + 0.3 * rng.integers(0, 256, (h, w, 3))
    return img.clip(0, 255).astype(np.uint8)


def checksum(sd):
    return float(sum(v.double().abs().sum().item() for k, v in sorted(sd.items()) if not k.startswith("__")))


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max())


def gen_clip():
    from semabs_b200.clip.model import synthetic_clip_state_dict
    from semabs_b200.clip.tokenizer import tokenize

    out = {}
    # ---------------- ViT-B/32 ----------------
    sd_raw = synthetic_clip_state_dict("ViT-B/32", seed=0)
    sd = clip_oracle.convert_weights_values(sd_raw)
    out["b32_weights_checksum"] = np.float64(checksum(sd_raw))
    wrapper = ref_import.make_reference_wrapper("ViT-B/32", sd_raw)
    gradcam = wrapper.clip_gradcam
    gradcam.templates = [PROMPT]
    gradcam.set_classes(LABELS4)
    W_ref = torch.cat([gradcam.class_to_language_feature[l] for l in LABELS4], dim=1)
    tokens = tokenize([PROMPT.format(l) for l in LABELS4])
    W_or = clip_oracle.zeroshot_weights(sd, tokens, 4, 1)
    print("B/32 zeroshot weights oracle-vs-reference", rel_err(W_or, W_ref))
    assert rel_err(W_or, W_ref) < 1e-5
    out["b32_tokens"] = tokens.numpy().astype(np.int32)
    out["b32_W"] = W_ref.numpy()

    g = torch.Generator().manual_seed(1)
    tiles = torch.randn(3, 3, 224, 224, generator=g)
    gradcam.positive_attn_only = True
    rel_ref = gradcam(x=tiles, o=LABELS4).detach()
    rel_or = clip_oracle.relevancy(sd, tiles, W_ref)
    print("B/32 relevancy oracle-vs-reference", rel_err(rel_or, rel_ref), tuple(rel_ref.shape))
    assert rel_err(rel_or, rel_ref) < 1e-5
    out["b32_rel"] = rel_ref.numpy()
    gradcam.positive_attn_only = False
    rel_ref2 = gradcam(x=tiles, o=LABELS4).detach()
    assert rel_err(clip_oracle.relevancy(sd, tiles, W_ref, positive_attn_only=False), rel_ref2) < 1e-5
    out["b32_rel_signed"] = rel_ref2.numpy()

    # full get_clip_saliency: 2-scale pyramid with flipping, no jitter (deterministic)
    img = synth_image(5, 96, 96)
    cfg = dict(distractor_labels={}, horizontal_flipping=True, augmentations=0, imagenet_prompt_ensemble=False,
               positive_attn_only=True,
               cropping_augmentations=[{"tile_size": 96, "stride": 24}, {"tile_size": 48, "stride": 12}])  # fmt: skip
    t0 = time.time()
    maps_ref, feats_ref = wrapper.get_clip_saliency(img=img, text_labels=np.array(LABELS4), prompts=[PROMPT], **cfg)
    print("B/32 get_clip_saliency reference: %.1fs" % (time.time() - t0), tuple(maps_ref.shape))
    maps_or = clip_oracle.get_clip_saliency(sd, img, W_ref, cfg["cropping_augmentations"], horizontal_flipping=True)
    print("B/32 saliency map oracle-vs-reference", rel_err(maps_or, maps_ref),
          "argmax equal:", bool((maps_or.flatten(1).argmax(1) == maps_ref.flatten(1).argmax(1)).all()))
    assert rel_err(maps_or, maps_ref) < 1e-4
    out["b32_maps"] = maps_ref.numpy()
    out["b32_text_feats"] = feats_ref.numpy()
    # "chefer_et_al" single-tile config on a non-square-friendly size
    cfg1 = dict(cfg, horizontal_flipping=False, cropping_augmentations=[{"tile_size": 96, "stride": 24}])
    maps1, _ = wrapper.get_clip_saliency(img=img, text_labels=np.array(LABELS4), prompts=[PROMPT], **cfg1)
    out["b32_maps_single"] = maps1.numpy()

    # ---------------- ViT-L/14 (257 tokens: positional quirk, 13 rollout blocks) ----------------
    sd_raw = synthetic_clip_state_dict("ViT-L/14", seed=0)
    sd = clip_oracle.convert_weights_values(sd_raw)
    out["l14_weights_checksum"] = np.float64(checksum(sd_raw))
    wrapper = ref_import.make_reference_wrapper("ViT-L/14", sd_raw)
    gradcam = wrapper.clip_gradcam
    gradcam.templates = [PROMPT]
    labels = LABELS4[:2]
    gradcam.set_classes(labels)
    W_ref = torch.cat([gradcam.class_to_language_feature[l] for l in labels], dim=1)
    tokens = tokenize([PROMPT.format(l) for l in labels])
    assert rel_err(clip_oracle.zeroshot_weights(sd, tokens, 2, 1), W_ref) < 1e-5
    g = torch.Generator().manual_seed(2)
    tiles = torch.randn(2, 3, 224, 224, generator=g)
    gradcam.positive_attn_only = True
    t0 = time.time()
    rel_ref = gradcam(x=tiles, o=labels).detach()
    print("L/14 reference relevancy: %.1fs" % (time.time() - t0))
    t0 = time.time()
    rel_or = clip_oracle.relevancy(sd, tiles, W_ref)
    print("L/14 oracle relevancy: %.1fs; oracle-vs-reference %.2e" % (time.time() - t0, rel_err(rel_or, rel_ref)))
    assert rel_err(rel_or, rel_ref) < 1e-5
    out["l14_tokens"] = tokens.numpy().astype(np.int32)
    out["l14_W"] = W_ref.numpy()
    out["l14_rel"] = rel_ref.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "clip_golden.npz"), **out)
    print("wrote clip_golden.npz")


def gen_config0():
    """BASELINE.json configs[0]: the reference's own documented case — `generate_relevancy.py image`
    on /root/reference/matterport.png (976^2), ViT-B/32, the first 4 typer default labels, the shipped
    `saliency_configs["chefer_et_al"](976)` (one 976-px tile, no jitter / flip).  The input image is committed next to the
    golden as a LOSSLESS re-encoding (tests/golden/matterport_976.webp, bit-identical pixels, checked below) because
    /root/reference does not exist on the GPU box.  Stored: the reference maps sub-sampled 8x (the full 4x976^2 fp32
    output is 15 MB), plus per-map peak index / peak value / float64 sum of the FULL maps; the GPU test additionally
    compares the full maps with the oracle, which this function pins to the reference at full size."""
    from PIL import Image

    from semabs_b200.clip.model import synthetic_clip_state_dict

    src = os.path.join(ref_import.REF_ROOT, "matterport.png")
    img = np.array(Image.open(src).convert("RGB"))
    assert img.shape == (976, 976, 3) and img.dtype == np.uint8
    webp = os.path.join(GOLDEN, "matterport_976.webp")
    Image.fromarray(img).save(webp, "WEBP", lossless=True, quality=100, method=6)
    assert np.array_equal(np.array(Image.open(webp).convert("RGB")), img), "lossless re-encoding changed pixels"
    sd_raw = synthetic_clip_state_dict("ViT-B/32", seed=0)
    sd = clip_oracle.convert_weights_values(sd_raw)
    wrapper = ref_import.make_reference_wrapper("ViT-B/32", sd_raw)
    cfg = wrapper_cfg = import_saliency_configs()["chefer_et_al"](976)
    t0 = time.time()
    maps_ref, feats_ref = wrapper.get_clip_saliency(img=img, text_labels=np.array(LABELS4), prompts=[PROMPT], **cfg)
    print("config[0] reference get_clip_saliency: %.2fs" % (time.time() - t0), tuple(maps_ref.shape))
    W_ref = torch.cat([wrapper.clip_gradcam.class_to_language_feature[l] for l in LABELS4], dim=1)
    maps_or = clip_oracle.get_clip_saliency(sd, img, W_ref, cfg["cropping_augmentations"], positive_attn_only=True)
    e = rel_err(maps_or, maps_ref)
    same_peak = bool((maps_or.flatten(1).argmax(1) == maps_ref.flatten(1).argmax(1)).all())
    print("config[0] oracle-vs-reference", e, "argmax equal:", same_peak)
    assert e < 1e-5 and same_peak
    flat = maps_ref.flatten(1)
    np.savez_compressed(os.path.join(GOLDEN, "config0_golden.npz"),
                        maps_sub8=maps_ref[:, ::8, ::8].numpy(), peak_index=flat.argmax(1).numpy(),
                        peak_value=flat.max(1).values.numpy(), map_sum=flat.double().sum(1).numpy(),
                        n_at_peak=(flat == flat.max(1, keepdim=True).values).sum(1).numpy(),
                        text_feats=feats_ref.numpy(), image_checksum=np.int64(img.astype(np.int64).sum()))
    print("wrote config0_golden.npz; pixels tied at the reference peak per map:", (flat == flat.max(1, keepdim=True).values).sum(1).tolist())


def import_saliency_configs():
    return ref_import.import_reference_clip().saliency_configs


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if what in ("clip", "all"):
        gen_clip()
    if what in ("config0", "all"):
        gen_config0()
    if what in ("unet", "net", "all"):
        from oracle import gen_golden_3d

        gen_golden_3d.main()
