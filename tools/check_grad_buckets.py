"""Two (or more) ranks, one GPU each, under torch.distributed.run: the per-level gradient buckets reduced DURING the UNet
backward (train.GradientBuckets) must give the gradients one flat all-reduce after backward gives, and the VOOL train
step of bench.py is timed both ways.  Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_grad_buckets.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from semabs_b200 import train  # noqa: E402
from semabs_b200.net import SemAbs3D  # noqa: E402

BOUNDS = ((-1.0, -1.0, -0.1), (1.0, 1.0, 1.9))


def gradients(dev, rank, mode):
    os.environ["SEMABS_GRAD_BUCKETS"] = mode
    torch.manual_seed(5)
    m = SemAbs3D(voxel_shape=(32, 32, 32), scene_bounds=BOUNDS, unet_num_channels=16, unet_f_maps=16, unet_num_groups=8,
                 unet_num_levels=4, network_inputs=["saliency"], use_pts_feat_extractor=True, pts_feat_extractor_hidden_dim=128,
                 reduce_method="max", device=str(dev), batch_size=1).to(dev)
    g = torch.Generator().manual_seed(50 + rank)  # a different scene on every rank
    lo, hi = torch.tensor(BOUNDS[0]), torch.tensor(BOUNDS[1])
    B, P, n_in, n_out = 1, 2, 4000, 6000
    batch = dict(input_xyz_pts=(lo + (hi - lo) * torch.rand(B, n_in, 3, generator=g)).to(dev),
                 input_feature_pts=torch.randn(B, P, n_in, 1, generator=g).to(dev), tsdf_vol=torch.ones(B, 1, device=dev),
                 output_xyz_pts=(lo + (hi - lo) * torch.rand(B, P, n_out, 3, generator=g)).to(dev),
                 output_label_pts=(torch.rand(B, P, n_out, generator=g) < 0.15).float().to(dev),
                 out_of_bounds_pts=torch.zeros(B, P, n_out, dtype=torch.bool, device=dev),
                 out_of_frustum_pts_mask=torch.zeros(B, P, n_out, dtype=torch.bool, device=dev), patch_labels=[("a",), ("b",)])
    stats, _ = train.get_losses_ovssc(m, batch)
    with train.GradientBuckets(keep_local=True) as gb:
        stats["loss"].backward()
    names = {p: n for n, p in m.named_parameters()}
    local = {names[p]: t for p, t in (gb.local or {}).items()}  # bucketed parameters: this rank's own gradient
    local.update({n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None and p not in gb.reduced})
    train.all_reduce_gradients(m.parameters(), skip=gb.reduced)
    return {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}, local, gb.n_buckets, len(gb.reduced)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    # The same backward pass, so no run-to-run spread enters (two runs of the backward differ by ~1e-5, and by ~2e-3 in the
    # rare run where one ReLU input sits within rounding noise of zero — tools/debug_bwd_determinism.py): the gradients
    # the step ends up with must be the mean over ranks of the gradients each rank computed.
    world_t = float(world)
    worst, worst_key = 0.0, ""
    counts = {}
    for mode in ("0", "1"):
        g_avg, local, nb, nr = gradients(dev, rank, mode)
        counts[mode] = (nb, nr, len(local) - nr)
        for k in sorted(g_avg):
            parts = [torch.empty_like(local[k]) for _ in range(world)]
            dist.all_gather(parts, local[k].contiguous())
            mean = torch.stack(parts).sum(0) / world_t
            d = ((g_avg[k] - mean).norm() / mean.norm().clamp_min(1e-30)).item()
            if d > worst:
                worst, worst_key = d, f"{k} (buckets {'on' if mode == '1' else 'off'})"
        if mode == "1":
            g_bkt = g_avg
    (nb0, nr0, _), (nb1, nr1, n_rest) = counts["0"], counts["1"]
    assert (nb0, nr0) == (0, 0) and nb1 > 1 and nr1 > 0, counts
    # every rank must hold the same averaged gradients
    sig = torch.stack([g_bkt[k].double().sum() for k in sorted(g_bkt)])
    all_sig = [torch.empty_like(sig) for _ in range(world)]
    dist.all_gather(all_sig, sig)
    same = all(torch.equal(all_sig[0], s) for s in all_sig)
    pk = bench.peaks()
    res = {}
    for mode in ("0", "1"):
        os.environ["SEMABS_GRAD_BUCKETS"] = mode
        res[mode] = bench.bench_train(dev, rank, world, pk, steps=3, warmup=2)
    if rank == 0:
        print(json.dumps({"world": world, "buckets": nb1, "bucketed_parameters": nr1, "flat_parameters": n_rest,
                          "worst_rel_diff_vs_mean_of_rank_gradients": worst, "worst_tensor": worst_key,
                          "ranks_identical": same,
                          "train_ms_flat": res["0"].get("ms_per_step"), "train_ms_buckets": res["1"].get("ms_per_step")}))
    assert worst <= (0.0 if world == 2 else 1e-6) and same
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
