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
