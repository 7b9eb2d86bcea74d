"""TEST INFRASTRUCTURE ONLY — numpy restatement of the geometry glue between the relevancy extractor and SemAbs3D
(reference point_cloud.py:8-67, visualize.py:283-298).  Pinned: `python -m oracle.pipeline_oracle` imports the unmodified
reference functions from /root/reference (point_cloud.get_pointcloud / filter_pts_bounds; visualize.get_sample_points is
restated from its 12 lines because importing visualize.py needs a dozen absent packages) and asserts exact agreement on
seeded inputs — last run in the build container: max |diff| = 0."""
from __future__ import annotations

import numpy as np


def transform_pointcloud(xyz_pts, rigid_transform):
    """point_cloud.py:8-21"""
    xyz = np.dot(rigid_transform[:3, :3], xyz_pts.T)
    xyz = xyz + np.tile(rigid_transform[:3, 3].reshape(3, 1), (1, xyz.shape[1]))
    return xyz.T


def get_pointcloud(depth_img, cam_intr, cam_pose=None):
    """point_cloud.py:34-67 (colour output dropped)"""
    h, w = depth_img.shape
    px, py = np.meshgrid(np.linspace(0, w - 1, w), np.linspace(0, h - 1, h))
    x = np.multiply(px - cam_intr[0, 2], depth_img / cam_intr[0, 0])
    y = np.multiply(py - cam_intr[1, 2], depth_img / cam_intr[1, 1])
    pts = np.array([x, y, depth_img]).transpose(1, 2, 0).reshape(-1, 3)
    if cam_pose is not None:
        pts = transform_pointcloud(pts, cam_pose)
    return pts


def filter_pts_bounds(xyz, bounds):
    """point_cloud.py:24-31"""
    m = xyz[:, 0] >= bounds[0, 0]
    for ax in range(3):
        m = np.logical_and(m, xyz[:, ax] >= bounds[0, ax])
        m = np.logical_and(m, xyz[:, ax] <= bounds[1, ax])
    return m


def get_sample_points(sampling_shape, scene_bounds):
    """visualize.py:283-298 (fp32 arithmetic like the torch original)"""
    idx = np.stack(np.meshgrid(*[np.arange(n) for n in sampling_shape], indexing="ij"), axis=-1).astype(np.float32)
    lc, uc = np.asarray(scene_bounds[0], np.float32), np.asarray(scene_bounds[1], np.float32)
    scales = (uc - lc) / (np.asarray(sampling_shape, np.float32) - 1)
    return (idx * scales + lc).reshape(-1, 3)


def _pin():
    import importlib.util, sys

    spec = importlib.util.spec_from_file_location("ref_point_cloud", "/root/reference/point_cloud.py")
    ref = importlib.util.module_from_spec(spec)
    for missing in ("pybullet", "pybullet_data", "matplotlib", "matplotlib.pyplot", "transforms3d", "skimage", "skimage.measure"):
        sys.modules.setdefault(missing, type(sys)(missing))  # imported at module level, unused by the pinned functions
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(0)
    depth = rng.uniform(0.3, 3.0, (37, 53)).astype(np.float32)
    K = np.array([[60.0, 0, 26.0], [0, 61.0, 18.0], [0, 0, 1]])
    T = np.array([[0.0, -1, 0, 0.2], [1, 0, 0, -0.1], [0, 0, 1, 0.4]])
    for pose in (None, T):
        a, b = ref.get_pointcloud(depth, None, K, pose)[0], get_pointcloud(depth, K, pose)
        assert np.abs(a - b).max() == 0
    bounds = np.array([[-1.0, -1.0, -0.1], [1.0, 1.0, 1.9]])
    pts = get_pointcloud(depth, K, T)
    assert (ref.filter_pts_bounds(pts, bounds) == filter_pts_bounds(pts, bounds)).all()
    print("pipeline oracle pinned to /root/reference/point_cloud.py: max |diff| = 0")


if __name__ == "__main__":
    _pin()


def prediction_volumes_reference(logits, sampling_shape, scene_bounds, depth, cam_intr, cam_extr, cutoff=-3.0):
    """TEST INFRASTRUCTURE: the tail of visualize.process_batch_ovssc (visualize.py:212-247) executed with the UNMODIFIED
    reference classes fusion.TSDFVolume and point_cloud.check_pts_in_frustum (imported from /root/reference; used only by
    the CPU pinning test, which is skipped where the checkout is absent)."""
    import importlib.util, sys

    import torch

    mods = {}
    for name in ("fusion", "point_cloud"):
        for missing in ("pybullet", "pybullet_data", "matplotlib", "matplotlib.pyplot", "transforms3d", "skimage", "skimage.measure"):
            sys.modules.setdefault(missing, type(sys)(missing))
        spec = importlib.util.spec_from_file_location("ref_" + name, f"/root/reference/{name}.py")
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[name] = m
    grid_points = get_sample_points(sampling_shape, scene_bounds)
    tsdf_vol = mods["fusion"].TSDFVolume(vol_bnds=np.array(scene_bounds).T,
                                         voxel_size=(scene_bounds[1][0] - scene_bounds[0][0]) / sampling_shape[0])
    tsdf_vol.integrate(color_im=np.zeros(depth.shape + (3,), np.uint8), depth_im=depth, cam_intr=cam_intr, cam_pose=cam_extr)
    # get_volume()[0] is this array; the colour half of get_volume overflows uint8 under numpy 2 (fusion.py:199-203)
    tsdf = tsdf_vol._tsdf_vol_cpu
    logprobs = torch.as_tensor(logits).permute(*range(1, logits.ndim), 0)
    prediction = logprobs.argmax(dim=-1)
    empty = (logprobs < cutoff).all(dim=-1).view(*sampling_shape)
    in_frustum = torch.from_numpy(mods["point_cloud"].check_pts_in_frustum(xyz_pts=grid_points, depth=depth, cam_pose=cam_extr,
                                                                          cam_intr=cam_intr)).view(*sampling_shape)
    out = []
    for c in range(logits.shape[0]):
        v = (prediction == c).float().view(*sampling_shape)
        v[empty] = 0.0
        v[~in_frustum] = 0.0
        v[torch.from_numpy(tsdf > 0.0)] = 0.0
        out.append(v)
    return torch.stack(out), tsdf
