"""CPU suite: the C-ABI library loads and exports every symbol include/semabs_b200.h declares; host-side pieces
(tokenizer, weight conversion, positional quirk) behave like the reference's."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "semabs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(semabs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge

    ge.build()
    from semabs_b200 import _lib

    lib = _lib.lib()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/semabs_b200.h but not exported"
    assert lib.semabs_abi_version() == 1


def test_no_cpu_fallback():
    from semabs_b200.unet3d import ResidualUNet3D

    m = ResidualUNet3D(in_channels=16, out_channels=16, f_maps=16, num_levels=2)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 16, 8, 8, 8))


def test_tokenizer_known_ids():
    from semabs_b200.clip.tokenizer import tokenize

    t = tokenize(["a photograph of a basketball jersey in a home."])
    assert t.shape == (1, 77)
    assert t[0, :12].tolist() == [49406, 320, 8853, 539, 320, 3835, 4471, 530, 320, 1137, 269, 49407]
    assert int(t[0].argmax()) == 11
    with pytest.raises(RuntimeError):
        tokenize(["word " * 100])


def test_convert_weights_and_positional_quirk():
    from oracle import clip_oracle
    from semabs_b200.clip import model

    sd = {"visual.proj": torch.randn(4, 4), "visual.ln_pre.weight": torch.randn(4), "x.mlp.c_fc.bias": torch.randn(4)}
    a, b = model.apply_convert_weights(sd), clip_oracle.convert_weights_values(sd)
    for k in sd:
        assert torch.equal(a[k], b[k])
    assert torch.equal(a["visual.ln_pre.weight"], sd["visual.ln_pre.weight"])
    assert torch.equal(a["visual.proj"], sd["visual.proj"].half().float())
    pos = torch.randn(50, 8)
    q = model.interpolate_positional_embedding(pos, 257)
    assert torch.equal(q, clip_oracle._positional_quirk(pos, 257))
    assert torch.equal(q[0], pos[0]) and q.shape == (257, 8)


def test_pillow_resize_tables():
    """Host side of semabs_tile_preprocess: the coefficient tables reproduce the installed Pillow's BICUBIC resize bit for
    bit when evaluated with Pillow's integer two-pass scheme (numpy emulation of what the kernel does)."""
    import numpy as np
    from PIL import Image

    from semabs_b200.clip.wrapper import pillow_bicubic_coeffs

    def resize(img, out):
        s = img.shape[0]
        kk, b = pillow_bicubic_coeffs(s, out)

        def one_pass(src):  # along axis 1
            dst = np.zeros((src.shape[0], out, 3), np.uint8)
            for xx in range(out):
                xmin, cnt = b[xx]
                acc = np.full((src.shape[0], 3), 1 << 21, np.int64)
                for x in range(cnt):
                    acc += src[:, xmin + x, :].astype(np.int64) * kk[xx, x]
                dst[:, xx, :] = np.clip(acc >> 22, 0, 255)
            return dst

        tmp = one_pass(img)                                   # horizontal
        return one_pass(tmp.transpose(1, 0, 2)).transpose(1, 0, 2)  # vertical

    rng = np.random.default_rng(0)
    for s in (336, 224, 168, 112, 84, 61, 500):
        img = rng.integers(0, 256, (s, s, 3), dtype=np.uint8)
        ref = np.array(Image.fromarray(img).resize((224, 224), Image.BICUBIC))
        assert np.array_equal(resize(img, 224), ref), s
