"""Relevancy store: device pack / unpack against the torch ops the reference calls
(F.interpolate nearest-exact at generate_relevancy.py:96-102, mean row :104-107, bilinear read-back dataset.py:866-871),
and the container round trip with the reference's key layout."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W", [(336, 336), (976, 976), (480, 640), (100, 37)])
def test_pack_matches_nearest_exact_and_mean(H, W):
    from semabs_b200 import relevancy_store as rs

    g = torch.Generator().manual_seed(H + W)
    maps = torch.rand(5, H, W, generator=g) * 0.01
    got = rs.pack_maps(maps.cuda()).cpu()
    ref = F.interpolate(maps[:, None], size=(128, 128), mode="nearest-exact")[:, 0]
    ref = torch.cat([ref, ref.mean(dim=0, keepdim=True)], dim=0)
    assert got.shape == (6, 128, 128)
    assert torch.equal(got[:5], ref[:5])  # a gather: bit-exact
    assert torch.allclose(got[5], ref[5], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("H,W", [(336, 336), (480, 640)])
def test_unpack_matches_reference_reader(H, W):
    from semabs_b200 import relevancy_store as rs

    g = torch.Generator().manual_seed(H)
    stored = torch.rand(9, 128, 128, generator=g) * 0.01
    rows, mean_row = [1, 3, 4, 7], 8
    got = rs.unpack_maps(stored.cuda(), rows, (H, W), mean_row, gain=50.0).cpu()
    ref = F.interpolate((stored[rows] - stored[mean_row])[:, None], size=(H, W), mode="bilinear", align_corners=False)[:, 0] * 50
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 1e-6 * ref.abs().max().item() + 1e-9
    plain = rs.unpack_maps(stored.cuda(), rows, (H, W), None).cpu()
    ref2 = F.interpolate(stored[rows][:, None], size=(H, W), mode="bilinear", align_corners=False)[:, 0]
    assert (plain - ref2).abs().max().item() < 1e-6 * ref2.abs().max().item()


def test_container_round_trip(tmp_path):
    from semabs_b200 import relevancy_store as rs

    g = torch.Generator().manual_seed(3)
    labels = ["chair", "lamp", "sofa"]
    maps = torch.rand(3, 336, 336, generator=g) * 0.01
    feats = torch.randn(3, 512, generator=g)
    path = str(tmp_path / "scene.npz")
    st = rs.RelevancyStore(path)
    refs = st.add("rgb", "ours", maps.cuda(), labels, feats)
    assert refs.tolist() == [0, 1, 2, 3]
    st.add("domain_randomized_rgb", "ours", (maps * 2).cuda(), labels, feats)
    with pytest.raises(Exception):
        st.add("rgb", "ours", maps.cuda(), labels, feats)  # already present (write_to_hdf5 semantics)
    st.flush()
    rd = rs.RelevancyStore(path)
    assert rd.arrays["saliencies"].shape == (8, 128, 128) and rd.arrays["saliencies"].dtype == np.float32
    assert rd.arrays["data/saliencies/rgb|ours|saliency_text_labels"].astype(str).tolist() == labels + ["mean"]
    tf = rd.arrays["data/saliencies/rgb|ours|saliency_text_label_features"]
    assert tf.shape == (4, 512) and np.allclose(np.linalg.norm(tf, axis=1), 1.0, atol=1e-5)
    out = rd.load_patches("rgb", "ours", (336, 336), labels=["sofa", "chair"], gain=50.0)
    assert out["patch_labels"].tolist() == ["chair", "sofa"]  # HDF5 indexing must be in order (dataset.py:811-812)
    packed = F.interpolate(maps[:, None], size=(128, 128), mode="nearest-exact")[:, 0]
    ref = F.interpolate((packed[[0, 2]] - packed.mean(0))[:, None], size=(336, 336), mode="bilinear", align_corners=False)[:, 0] * 50
    assert (out["patch_saliencies"].cpu() - ref).abs().max().item() < 1e-5 * ref.abs().max().item()
    second = rd.load_patches("domain_randomized_rgb", "ours", (336, 336), subtract_mean_relevancy=False)
    assert second["num_patches"] == 3 and second["patch_saliencies"].shape == (3, 336, 336)
