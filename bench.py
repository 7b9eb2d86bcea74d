"""Benchmark of the hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One step = one pass of the relevancy hot path over one batch of synthetic input per GPU:
BASELINE.json configs[1] — 8 images 336x336 RGB, ViT-L/14 (seeded random-init weights: no checkpoints offline),
5-size crop pyramid (285 tiles / image, no jitter / flip), 16 labels -> 128 relevancy maps [336,336] per step per GPU.
`value` = relevancy maps / s with the preprocessed tiles already resident in HBM; `e2e` = the same metric through
ClipWrapper (host uint8 images -> PIL tile preprocessing -> H2D -> kernels -> D2H of the fp32 maps).
The second half of BASELINE.json's metric (voxel grids / s, configs[2]: ResidualUNet3D 128^3 x 32 ch, batch 4) is
measured in the same run and reported under "voxel".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

LABELS16 = ["basketball jersey", "nintendo switch", "television", "ping pong table", "vase", "fireplace",
            "abstract painting of a vespa", "carpet", "wall", "microwave", "cabinet", "fire extinguisher", "mirror",
            "woven chair", "globe", "leather sofa"]  # fmt: skip
PROMPT = "a photograph of a {} in a home."
PYRAMID = [{"tile_size": s, "stride": s // 4} for s in (336, 224, 168, 112, 84)]  # 1+9+25+81+169 = 285 tiles
IMG = 336
IMAGES_PER_STEP = 8
# tiles per engine call: the reference's `tile_batch_size` kwarg (a memory knob there, default 32, no effect on results);
# 95 = 285 / 3 keeps every GEMM's M dimension large and leaves no ragged last batch
TILE_BATCH = int(os.environ.get("SEMABS_TILE_BATCH", "95"))  # tiles per engine pass (285 = 3 x 95); env override for A/B runs
MODEL = "ViT-L/14"
# algorithmic work (SURVEY.md §8d / BASELINE.md §3): per tile 162.0 GF forward + 89.1 GF backward per label
GF_FWD_TILE, GF_BWD_TILE_LABEL = 162.0, 89.1
UNET_GF_PER_GRID = {16: 340.8, 32: 1363.2}
UNET_GB_PER_GRID_FP32 = {16: 3.35, 32: 6.71}


def synth_image(seed):
    return np.random.default_rng(seed).integers(0, 256, (IMG, IMG, 3), dtype=np.uint8)


_REAL_STDOUT = None


def _quiet_stdout():
    """The contract is ONE JSON line on stdout; libraries (NCCL prints its version there when NCCL_DEBUG is set) write to
    fd 1 behind Python's back, so fd 1 is pointed at stderr for the run and the line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def _release():
    """Drop dead modules / tapes / workspaces (autograd nodes and modules reference each other) and return the memory."""
    import gc

    gc.collect()
    torch.cuda.empty_cache()


def measured_traffic():
    """DRAM bytes from the committed ncu captures (profiles/rNN_traffic.json of the latest round, made with the commands in its
    `source` fields); bench.py itself never runs under a profiler."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return json.load(open(p))
    return {}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline / CUDA-eager comparator: the STOCK reference code path, executed from the byte-for-byte
# snapshot oracle/_ref (oracle/build_ref.py; /root/reference itself does not exist on the GPU box) with the same seeded
# weights as our arm.  Falls back to the oracle restatement (kind "port") only when the snapshot is missing.
# ---------------------------------------------------------------------------------------------------------
_STOCK = {}


def stock_reference():
    """-> oracle.ref_import bound to the snapshot, or None"""
    from oracle import build_ref

    if not build_ref.available():
        return None
    os.environ["SEMABS_REFERENCE_ROOT"] = build_ref.DST
    from oracle import ref_import

    return ref_import


def stock_wrapper(device):
    """the reference's ClipWrapper singleton (CLIP/clip/__init__.py:44-101) on `device`, seeded ViT-L/14 weights"""
    key = str(device)
    if _STOCK.get("key") != key:
        from semabs_b200.clip.model import synthetic_clip_state_dict

        ri = stock_reference()
        _STOCK["wrapper"] = ri.make_reference_wrapper(MODEL, synthetic_clip_state_dict(MODEL, seed=0), device=device)
        _STOCK["key"] = key
    return _STOCK["wrapper"]


STOCK_CFG = dict(distractor_labels={}, horizontal_flipping=False, augmentations=0, positive_attn_only=True)


def cpu_relevancy_sample(n_labels=16):
    """Times the reference's own CPU path on a BOUNDED sample of configs[1]: stock ClipWrapper.get_clip_saliency (fp32 on
    CPU, all host threads) on a 336^2 synthetic image with 16 labels and a ONE-tile pyramid, plus its text tower alone;
    a full image costs t_text + 285 x (t_call - t_text) (every tile is resized to 224^2, so tiles cost the same)."""
    torch.set_num_threads(os.cpu_count())
    img = synth_image(0)
    labels = LABELS16[:n_labels]
    n_full = 285
    if stock_reference() is not None:
        w = stock_wrapper("cpu")
        t0 = time.perf_counter()
        w.clip_gradcam.templates = [PROMPT]
        w.clip_gradcam.set_classes(labels)
        t_text = time.perf_counter() - t0
        t0 = time.perf_counter()
        maps, _ = w.get_clip_saliency(img=img, text_labels=np.array(labels), prompts=[PROMPT],
                                      cropping_augmentations=[{"tile_size": IMG, "stride": IMG // 4}], **STOCK_CFG)
        t_call = time.perf_counter() - t0
        assert tuple(maps.shape) == (n_labels, IMG, IMG)
        t_image = t_text + n_full * max(t_call - t_text, 1e-9)
        kind = "reference"
        sample = (f"stock ClipWrapper.get_clip_saliency from the oracle/_ref snapshot: 1 tile x {n_labels} labels of the 285-tile "
                  f"pyramid ({MODEL}, fp32, torch CPU, {torch.get_num_threads()} threads) in {t_call:.1f} s incl. {t_text:.1f} s text tower; "
                  f"per image = text + 285 x tile")
    else:
        from oracle import clip_oracle
        from semabs_b200.clip.model import synthetic_clip_state_dict

        sd = clip_oracle.convert_weights_values(synthetic_clip_state_dict(MODEL, seed=0))
        desc = clip_oracle.enumerate_tiles(img.shape, PYRAMID)
        tiles = torch.stack([clip_oracle.preprocess_tile(img[r : r + s, c : c + s]) for r, c, s in desc[[0, len(desc) - 1]]])
        g = torch.Generator().manual_seed(0)
        W = torch.randn(768, n_labels, generator=g)
        W = (W / W.norm(dim=0, keepdim=True)).contiguous()
        t0 = time.perf_counter()
        clip_oracle.relevancy(sd, tiles, W)
        t_call = time.perf_counter() - t0
        t_image = n_full * t_call / 2
        kind = "port"
        sample = (f"oracle restatement (faster than the stock path: one autograd call per label instead of 13): 2 tiles x {n_labels} "
                  f"labels in {t_call:.1f} s, extrapolated linearly in tile count")
    return {"value": n_labels / t_image, "unit": "relevancy-maps/s", "cores": os.cpu_count(), "kind": kind, "sample": sample,
            "sample_seconds": t_call}


def _stock_unet(C, device):
    ri = stock_reference()
    torch.manual_seed(0)
    if ri is not None:
        m = ri.import_reference_module("unet3d").ResidualUNet3D(in_channels=C, out_channels=C, f_maps=C, num_groups=8, num_levels=6)
        return m.to(device).eval(), "reference"
    from oracle import unet_oracle
    from semabs_b200.unet3d import ResidualUNet3D

    sd = {k: v.detach().to(device) for k, v in ResidualUNet3D(in_channels=C, out_channels=C, f_maps=C, num_groups=8, num_levels=6).state_dict().items()}
    return (lambda x: unet_oracle.residual_unet3d(sd, x)), "port"


def cpu_voxel_sample(C=32):
    """The voxel half on the host: the stock ResidualUNet3D (reference unet3d.py from the snapshot, fp32, torch CPU, all
    threads) on ONE 128^3 x C grid — the same architecture / sizes as bench_voxel."""
    torch.set_num_threads(os.cpu_count())
    m, kind = _stock_unet(C, "cpu")
    x = torch.randn(1, C, 128, 128, 128, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        t0 = time.perf_counter()
        m(x)
        dt = time.perf_counter() - t0
    return {"value": 1.0 / dt, "unit": "voxel-grids/s", "cores": os.cpu_count(), "kind": kind,
            "sample": f"1 grid 128^3 x {C} ch through the 6-level ResidualUNet3D (fp32, torch CPU, {torch.get_num_threads()} threads) in {dt:.1f} s"}


def cuda_eager_relevancy(dev):
    """SURVEY.md §8d's practical GPU comparator: the STOCK reference Python on torch CUDA eager on this B200 (fp16 model, as
    `load` leaves it on CUDA), same image / labels / pyramid as configs[1].  A 10-tile sub-pyramid is timed first; the full
    285-tile image is run when that predicts under ~3 minutes, otherwise the sample is extrapolated by tile count."""
    if stock_reference() is None:
        return {"unavailable": "oracle/_ref snapshot missing"}
    try:
        w = stock_wrapper(dev)
        img = synth_image(0)
        call = lambda pyr: w.get_clip_saliency(img=img, text_labels=np.array(LABELS16), prompts=[PROMPT],
                                               cropping_augmentations=pyr, **STOCK_CFG)[0]
        call(PYRAMID[:1])  # warm-up (cuBLAS / cuDNN handles, allocator)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        call(PYRAMID[:2])
        torch.cuda.synchronize()
        t10 = time.perf_counter() - t0
        if t10 / 10 * 285 < 180:
            t0 = time.perf_counter()
            maps = call(PYRAMID)
            torch.cuda.synchronize()
            t = time.perf_counter() - t0
            note = f"one full image (285 tiles x 16 labels, tile_batch_size 32 = the reference default) in {t:.1f} s"
        else:
            t = t10 / 10 * 285
            note = f"10-tile sub-pyramid x 16 labels in {t10:.1f} s, extrapolated by tile count to 285"
        return {"value": len(LABELS16) / t, "unit": "relevancy-maps/s", "impl": "stock reference (oracle/_ref) on torch CUDA eager, fp16 model",
                "sample": note, "torch": torch.__version__}
    except Exception as e:  # the stock code pre-dates torch 2: report instead of hiding
        return {"unavailable": f"stock reference failed on CUDA: {type(e).__name__}: {str(e)[:200]}"}
    finally:
        _STOCK.clear()
        _release()


@torch.no_grad()
def cuda_eager_voxel(dev, C=32, N=4):
    """stock ResidualUNet3D on torch CUDA eager (cuDNN conv3d), fp32 with TF32 off (the accuracy class of our precise mode)
    and on (torch's default for convolutions), batch 4 x 128^3 x 32 ch."""
    out = {}
    try:
        m, kind = _stock_unet(C, dev)
        x = torch.randn(N, C, 128, 128, 128, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
        old = torch.backends.cudnn.allow_tf32
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            for _ in range(2):
                m(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                m(x)
            e1.record()
            torch.cuda.synchronize()
            out[name] = {"value": 3 * N / (e0.elapsed_time(e1) / 1e3), "unit": "voxel-grids/s", "ms_per_batch": e0.elapsed_time(e1) / 3}
        torch.backends.cudnn.allow_tf32 = old
        out["impl"] = f"stock unet3d.ResidualUNet3D ({kind}) on torch {torch.__version__} CUDA eager, cuDNN conv3d, batch {N}"
        del m, x
    except Exception as e:
        out["unavailable"] = f"{type(e).__name__}: {str(e)[:200]}"
    _release()
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on the box's host cores, all threads, on our arm's config /
    metric / unit; each step = the bounded sample described in cpu_baseline.sample.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(min(max(args.warmup, 0), 1)):
        cpu_relevancy_sample(2)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = cpu_relevancy_sample(16)
        vals.append(r["value"])
    dt = time.perf_counter() - t0
    v = float(np.mean(vals))
    r["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v,
            "unit": "relevancy-maps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(IMAGES_PER_STEP, 285),
            "cpu_baseline": r, "e2e": {"value": v, "unit": "relevancy-maps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "voxel": dict(cpu_voxel_sample(), metric="voxel-grids/sec/GPU (128^3, 32 ch)")}
    emit(line)


METRIC = "relevancy-maps/sec/GPU (336^2, 5 scales, 16 labels)"


def workload_config(images, n_tiles):
    """identical in both arms (the reference arm runs a bounded sample OF this workload, described in its cpu_baseline)"""
    return {"workload": f"configs[1]: {MODEL} (seeded random init), {images} images 336x336 per GPU per step, {n_tiles} tiles/image "
                        f"(5 crop sizes), {len(LABELS16)} labels, no jitter/flip"}


def family_table(prof: dict, pk: dict):
    """CALL_PROFILE records -> [{kernel, ms, share, calls, bound, achieved, unit, frac}] sorted by time"""
    total = sum(v["ms"] for v in prof.values()) or 1.0
    rows = []
    for name, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        row = {"kernel": name, "ms": round(v["ms"], 3), "share": round(v["ms"] / total, 4), "calls": v["calls"]}
        if v.get("flops"):
            tf = v["flops"] / (v["ms"] * 1e-3) / 1e12
            row.update(bound="tensor", achieved=round(tf, 1), unit="TFLOP/s", frac=round(tf / pk["tflops"], 4))
        elif v.get("bytes"):
            gb = v["bytes"] / (v["ms"] * 1e-3) / 1e9
            row.update(bound="hbm", achieved=round(gb, 1), unit="GB/s", frac=round(gb / pk["hbm_gbs"], 4))
        rows.append(row)
    return rows


# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=IMAGES_PER_STEP)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-voxel", action="store_true")
    ap.add_argument("--skip-train", action="store_true")
    ap.add_argument("--skip-pipeline", action="store_true")
    ap.add_argument("--skip-amp", action="store_true")
    ap.add_argument("--skip-eager", action="store_true")
    ap.add_argument("--skip-ours", action="store_true")
    ap.add_argument("--train-descs", type=int, default=16)
    args = ap.parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from semabs_b200 import ops
    from semabs_b200.clip import ClipWrapper
    from semabs_b200.clip.tokenizer import tokenize

    ClipWrapper.reset()
    ClipWrapper(MODEL, dev, seed=0)
    gc = ClipWrapper.clip_gradcam
    eng = gc.engine
    gc.templates = [PROMPT]
    gc.set_classes(LABELS16)
    cfg = dict(augmentations=0, cropping_augmentations=PYRAMID)
    imgs = [synth_image(rank * 1000 + i) for i in range(args.images)]
    P = len(LABELS16)

    # ---- device-resident inputs for `value` ----
    pre = [ClipWrapper.create_tiles(img=im, **cfg) for im in imgs]
    dev_tiles = [t.to(dev) for (_, t, _) in pre]
    n_tiles = pre[0][0].shape[0]

    def step_device():
        out = None
        for (desc, _, order), tl in zip(pre, dev_tiles):
            out = ClipWrapper.get_clip_saliency_device(tl, desc, order, LABELS16, IMG, IMG, positive_attn_only=True,
                                                       tile_batch_size=TILE_BATCH)
        return out

    def step_e2e():
        """the reference's public call, whole (CLIP/clip/__init__.py:103-133): set_classes (tokeniser + text tower) ->
        tile pyramid -> relevancy -> assembly -> fp32 maps on the host, per image"""
        last = None
        for im in imgs:
            maps, feats = ClipWrapper.get_clip_saliency(img=im, text_labels=np.array(LABELS16), prompts=[PROMPT], distractor_labels={},
                                                        horizontal_flipping=False, positive_attn_only=True,
                                                        tile_batch_size=TILE_BATCH, **cfg)
            last = maps
        return last

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item()

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.GEMM_PROFILE.enable()
    launches0 = eng.kernel_launches
    ms_dev, _ = timed(step_device, args.steps)
    launches = eng.kernel_launches - launches0 + args.steps * args.images  # + one assembly kernel per image
    gemm_prof = ops.GEMM_PROFILE.collect()
    clocks = sampler.stop() if rank == 0 else None
    maps_per_step = args.images * P
    value = world * maps_per_step * args.steps / (ms_dev / 1e3)

    # ---- per-kernel-family roofline table: ONE image (285 tiles x 16 labels) with a CUDA-event bracket around every C-ABI
    # call, outside the timed region ----
    from semabs_b200._lib import CALL_PROFILE

    CALL_PROFILE.enable()
    desc0, _, order0 = pre[0]
    ClipWrapper.get_clip_saliency_device(dev_tiles[0], desc0, order0, LABELS16, IMG, IMG, positive_attn_only=True, tile_batch_size=TILE_BATCH)
    kernel_table = family_table(CALL_PROFILE.collect(), peaks())

    # ---- e2e through the public API (host images, PIL preprocessing, H2D, D2H) ----
    step_e2e()
    _, wall_e2e = timed(step_e2e, max(1, args.steps // 2))
    e2e_steps = max(1, args.steps // 2)
    e2e_value = world * maps_per_step * e2e_steps / (wall_e2e / 1e3)
    # with device tile preprocessing only the uint8 image and the tile table cross PCIe; the host path ships fp32 tiles
    h2d = args.images * (IMG * IMG * 3 + n_tiles * 5 * 4) if ClipWrapper.device_preprocessing else args.images * n_tiles * 3 * 224 * 224 * 4
    d2h = args.images * (P * IMG * IMG * 4 + P * 768 * 4)

    # ---- R2' (SURVEY.md §8d): the reference-faithful `saliency_configs["ours"](336)` TTA — 204 tiles x (1 + 5 ColorJitter
    # copies) x 2 flips = 2448 ViT passes per image — through the public call, one image (reported for fidelity, not the target)
    faithful = None
    if not args.skip_ours:
        from semabs_b200.clip import saliency_configs

        ours_cfg = saliency_configs["ours"](IMG)
        torch.manual_seed(0)
        call = lambda: ClipWrapper.get_clip_saliency(img=imgs[0], text_labels=np.array(LABELS16), prompts=[PROMPT],
                                                     tile_batch_size=102, **ours_cfg)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        call()
        torch.cuda.synchronize()
        t_ours = time.perf_counter() - t0
        faithful = {"metric": "relevancy-maps/s, saliency_configs['ours'](336): 204 tiles x 6 image copies (device ColorJitter) x 2 flips = 2448 passes/image",
                    "value": world * P / t_ours, "unit": "relevancy-maps/s", "s_per_image": t_ours,
                    "algorithmic_tflops": world * 2448 * (GF_FWD_TILE + P * GF_BWD_TILE_LABEL) / 1e3 / t_ours}

    pk = peaks()
    flops_per_map = (GF_FWD_TILE + P * GF_BWD_TILE_LABEL) * 1e9 * n_tiles / P
    tr = measured_traffic()
    roofline = {"bound": "tensor", "kernel": "gemm_f16_tn_kernel (tcgen05)", "achieved": gemm_prof["tflops"], "peak": pk["tflops"],
                "unit": "TFLOP/s", "frac": gemm_prof["tflops"] / pk["tflops"], "traffic": tr.get("gemm_dram_bytes_per_launch"),
                "traffic_note": tr.get("gemm_source"),
                "peak_source": pk["src"] + " (sustained bf16 cuBLAS)", "gemm_launches": gemm_prof["launches"],
                "gemm_time_share_of_step": gemm_prof["ms"] / ms_dev,
                "whole_path_algorithmic_tflops": value / world * flops_per_map / 1e12,
                "whole_path_frac": value / world * flops_per_map / 1e12 / pk["tflops"],
                "kernels": kernel_table,
                "kernels_note": "per C-ABI entry point (= kernel family; semabs_attn_bwd_tc = row + column + tail kernels), one image, "
                                "CUDA events around every call; achieved = algorithmic FLOPs or bytes of the calls / their summed time"}

    voxel = None
    if not args.skip_voxel:
        voxel = bench_voxel(dev, world, dist if world > 1 else None, pk)
    cuda_eager = None
    if world == 1 and not args.skip_eager:  # comparator legs: N = 1 only, like cpu_baseline
        cuda_eager = {"relevancy": cuda_eager_relevancy(dev), "voxel": cuda_eager_voxel(dev),
                      "note": "reference Python on torch CUDA eager on this GPU (SURVEY.md §8d), timed by wall clock around synchronised calls"}

    pipe = None
    if not args.skip_pipeline:
        pipe = bench_pipeline(dev, rank, world, cfg, imgs)  # configs[4]: 64 images sharded over 8 GPUs = 8 per rank
    # the training step keeps ~80 GB of activations: release everything the earlier sections hold first
    splits = (eng.fwd_splits, eng.bwd_splits)
    del dev_tiles, pre, eng, gc
    ClipWrapper.reset()
    _release()
    train = None
    if not args.skip_train:
        train = bench_train(dev, rank, world, pk, num_descs=args.train_descs)
        _release()
        if not args.skip_amp:
            train["amp_like"] = bench_train(dev, rank, world, pk, num_descs=args.train_descs, precise=False)
            _release()

    if rank == 0:
        skip_cpu = args.skip_cpu or world > 1  # reported baseline: rank 0 at N = 1 only
        cpu = None if skip_cpu else cpu_relevancy_sample(16)
        if voxel is not None and not skip_cpu:
            voxel["cpu_baseline"] = cpu_voxel_sample()
        line = {"metric": METRIC, "value": value, "unit": "relevancy-maps/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 MMA / f32 accumulate (fwd hi+lo split)",
                "data": "synthetic",
                "config": dict(workload_config(args.images, n_tiles),
                               l2_policy="inputs larger than L2: per-step working set ~6 GB of saved activations + 1.4 GB tiles",
                               tile_batch_size=TILE_BATCH, fwd_splits=splits[0], bwd_splits=splits[1]),
                "e2e": {"value": e2e_value, "unit": "relevancy-maps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "note": "ClipWrapper.get_clip_saliency per host uint8 image (the reference's public call, whole): tokeniser + text tower "
                                "(set_classes) -> tile crop / Pillow-exact bicubic / normalise on the GPU -> relevancy -> assembly -> D2H of the "
                                "fp32 maps and the text features"},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "cuda_eager": cuda_eager, "faithful_ours": faithful, "clocks": clocks,
                "voxel": voxel, "pipeline": pipe, "train": train}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


@torch.no_grad()
def bench_voxel(dev, world, dist, pk, C=32, N=4, steps=5, warmup=3):
    """configs[2]: ResidualUNet3D inference, 128^3 x 32 channels, batch 4 (6 levels, 8 groups)."""
    from semabs_b200.unet3d import ResidualUNet3D

    torch.manual_seed(0)
    m = ResidualUNet3D(in_channels=C, out_channels=C, f_maps=C, num_groups=8, num_levels=6).to(dev)
    x = torch.randn(N, C, 128, 128, 128, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
    for _ in range(warmup):
        m(x)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = m.kernel_launches
    e0.record()
    for _ in range(steps):
        y = m(x)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    # e2e: host NCDHW fp32 in pinned memory -> device -> forward -> pinned host buffer, EVERY batch; the three stages run on
    # three streams (copy-in of batch i+1 and copy-out of batch i-1 overlap the kernels of batch i), so the rate is set by
    # the slowest stage — 1 GiB over PCIe each way per 4 grids
    xh = x.cpu().pin_memory()
    yh = [torch.empty(N, C, 128, 128, 128).pin_memory() for _ in range(2)]
    xd = [torch.empty_like(x) for _ in range(2)]
    s_in, s_out, cur = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [None, None]    # compute has consumed xd[b]
    ev_out = [None, None]     # yh[b] has been written
    K = max(6, 4 * steps)  # the timed region includes the pipeline's fill (first copy-in) and drain (last copy-out): ~2 stages of ~20 ms

    def run_e2e(k):
        for i in range(k):
            b = i % 2
            with torch.cuda.stream(s_in):
                if ev_free[b] is not None:
                    s_in.wait_event(ev_free[b])
                xd[b].copy_(xh, non_blocking=True)
                ev_in[b].record(s_in)
            cur.wait_event(ev_in[b])
            y = m(xd[b])
            ev_free[b] = torch.cuda.Event()
            ev_free[b].record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_free[b])
                if ev_out[b] is not None:
                    ev_out[b].synchronize()  # (host-side reuse guard of the pinned output buffer)
                yh[b].copy_(y, non_blocking=True)
                y.record_stream(s_out)
                ev_out[b] = torch.cuda.Event()
                ev_out[b].record(s_out)
        torch.cuda.synchronize()

    run_e2e(6)  # warm-up: both input buffers seen twice (the second sight of a buffer captures its CUDA graph)
    t0 = time.perf_counter()
    run_e2e(K)
    e2e = K * N / (time.perf_counter() - t0) * world
    from semabs_b200._lib import CALL_PROFILE

    CALL_PROFILE.enable()
    m(x)
    kernel_table = family_table(CALL_PROFILE.collect(), pk)
    grids_s = world * N * steps / (ms / 1e3)
    per_gpu = grids_s / world
    precise_mode, voxel_launches = m.precise, m.kernel_launches - l0
    del m, xd, yh, xh, y
    _release()
    return {"metric": "voxel-grids/sec/GPU (128^3, 32 ch)", "value": grids_s, "unit": "voxel-grids/s", "ms_per_step": ms / steps,
            "batch": N, "precise": precise_mode, "gpu_launches": voxel_launches,
            "e2e": {"value": e2e, "unit": "voxel-grids/s", "h2d_bytes_per_step": N * C * 128**3 * 4, "d2h_bytes_per_step": N * C * 128**3 * 4},
            "roofline": {"bound": "hbm", "achieved": per_gpu * UNET_GB_PER_GRID_FP32[C], "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": per_gpu * UNET_GB_PER_GRID_FP32[C] / pk["hbm_gbs"],
                         "traffic": (measured_traffic().get("unet_forward_dram_bytes") if C == 32 and N == 4 else None),
                         "algorithmic_bytes": N * UNET_GB_PER_GRID_FP32[C] * 1e9,
                         "tensor_tflops": per_gpu * UNET_GF_PER_GRID[C] / 1e3, "tensor_frac": per_gpu * UNET_GF_PER_GRID[C] / 1e3 / pk["tflops"],
                         "note": "whole-forward algorithmic bytes (BASELINE.md byte rule, fp32 I/O) / time",
                         "kernels": kernel_table}}


@torch.no_grad()
def bench_pipeline(dev, rank, world, cfg, images):
    """configs[4]: synthetic RGB-D -> 16-label relevancy (x50, minus label mean) -> depth back-projection -> SemAbs3D
    (reference defaults: 128^3 grid, 16 channels, 6 levels, 80k input points per class) -> logits on a 128^3 query
    lattice + class arg-max, device resident (semabs_b200.pipeline).  Timed per image through the public pipeline call
    (host uint8 image and fp32 depth in, device logits out + D2H of the int64 arg-max volume)."""
    import torch.distributed as dist

    from semabs_b200 import pipeline
    from semabs_b200.net import SemAbs3D

    bounds = ((-1.0, -1.0, -0.1), (1.0, 1.0, 1.9))
    torch.manual_seed(1)
    net = SemAbs3D(voxel_shape=(128, 128, 128), scene_bounds=bounds, unet_num_channels=16, unet_f_maps=16, unet_num_groups=8,
                   unet_num_levels=6, network_inputs=["saliency"], use_pts_feat_extractor=True,
                   pts_feat_extractor_hidden_dim=128, reduce_method="max", device=str(dev), batch_size=1).to(dev)
    rng = np.random.default_rng(5 + rank)
    # a tilted synthetic depth plane seen by a camera 1.2 m in front of the scene volume
    yy, xx = np.mgrid[0:IMG, 0:IMG].astype(np.float32)
    depth = (1.0 + 0.6 * xx / IMG + 0.3 * yy / IMG + 0.02 * rng.standard_normal((IMG, IMG))).astype(np.float32)
    K = np.array([[0.9 * IMG, 0, IMG / 2 - 0.5], [0, 0.9 * IMG, IMG / 2 - 0.5], [0, 0, 1]])
    T = np.array([[1.0, 0, 0, 0.0], [0, 0, 1, -1.2], [0, -1, 0, 0.9]])
    gen = torch.Generator(device=dev).manual_seed(3)
    lattice = (240, 240, 240)  # the reference's dense-sweep lattice (visualize.py:163-164)
    run = lambda im: pipeline.rgbd_to_ovssc_logits(net, im, depth, K, T, LABELS16, bounds, dict(positive_attn_only=True,
                                                   tile_batch_size=TILE_BATCH, **cfg), sampling_shape=lattice,
                                                   generator=gen)["prediction"].cpu()
    run(images[0])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for im in images:
        pred = run(im)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    n = len(images)
    del net
    _release()
    return {"metric": "RGB-D images/s through relevancy -> OVSSC logits (336^2, 16 labels, 128^3 grid, 240^3 query lattice)",
            "images_per_gpu": n, "workload": "configs[4]: 64 synthetic RGB-D images sharded over 8 GPUs = 8 per GPU (weak scaling: every N runs 8 per rank)",
            "value": world * n / dt.item(), "unit": "images/s (sum over GPUs)", "s_per_image": dt.item() / n,
            "relevancy_maps_per_s": world * n * len(LABELS16) / dt.item(), "classes_in_prediction": int(pred.unique().numel()),
            "h2d_bytes_per_image": IMG * IMG * 3 + IMG * IMG * 4 + 285 * 5 * 4, "d2h_bytes_per_image": 240**3 * 8}


def bench_train(dev, rank, world, pk, C=16, num_descs=16, steps=2, warmup=2, precise=True):
    """configs[3]: SemAbsVOOL train step with the reference's defaults (utils.py:38-77: 128^3 grid, 16 channels, 6 levels,
    80k input / 400k output points, pointing_dim 64, decoder_concat_xyz_pts, LAMB lr 1e-3 wd 1e-5, grad_max_norm 2.0),
    batch 1 scene x `num_descs` descriptions per GPU = 2 x num_descs voxel grids through the UNet forward + backward,
    gradient all-reduce over NCCL when world > 1."""
    import torch.distributed as dist

    from semabs_b200 import train
    from semabs_b200.net import SemAbsVOOL

    bounds = ((-1.0, -1.0, -0.1), (1.0, 1.0, 1.9))
    torch.manual_seed(0)  # same initial weights on every rank (what DDP's broadcast would give)
    net = SemAbsVOOL(pointing_method="cosine_sim", pointing_dim=64, device=str(dev), decoder_concat_xyz_pts=True,
                     voxel_shape=(128, 128, 128), scene_bounds=bounds, unet_num_channels=C, unet_f_maps=C, unet_num_groups=8,
                     unet_num_levels=6, network_inputs=["saliency"], use_pts_feat_extractor=True,
                     pts_feat_extractor_hidden_dim=128, reduce_method="max", batch_size=1, precise=precise).to(dev)
    opt = train.Lamb(net.parameters(), lr=1e-3, weight_decay=1e-5)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    lo, hi = torch.tensor(bounds[0], device=dev), torch.tensor(bounds[1], device=dev)
    n_in, n_out, B = 80000, 400000, 1
    rels = SemAbsVOOL.RELATIONS[:6]
    batch = dict(
        input_xyz_pts=lo + (hi - lo) * torch.rand(B, n_in, 3, device=dev, generator=g),
        input_target_saliency_pts=torch.randn(B, num_descs, n_in, 1, device=dev, generator=g),
        input_reference_saliency_pts=torch.randn(B, num_descs, n_in, 1, device=dev, generator=g),
        tsdf_vol=torch.ones(B, 1, device=dev),
        output_xyz_pts=lo + (hi - lo) * torch.rand(B, num_descs, n_out, 3, device=dev, generator=g),
        output_label_pts=(torch.rand(B, num_descs, n_out, device=dev, generator=g) < 0.1).float(),
        out_of_bounds_pts=torch.zeros(B, num_descs, n_out, dtype=torch.bool, device=dev),
        spatial_relation_name=[[rels[d % 6]] * B for d in range(num_descs)],
    )
    unet = net.completion_net.vol_feature_extractor
    torch.cuda.reset_peak_memory_stats(dev)
    losses = []
    for _ in range(warmup):
        losses.append(train.train_step(net, batch, train.get_losses_vool, opt, grad_max_norm=2.0)["loss"])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = unet.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        losses.append(train.train_step(net, batch, train.get_losses_vool, opt, grad_max_norm=2.0)["loss"])
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item() / steps
    peak_gb = torch.cuda.max_memory_allocated(dev) / 2**30
    launches_per_step = (unet.kernel_launches - l0) // steps
    loss_values = [float(x.detach()) for x in losses]
    del losses, opt, net, unet, batch
    _release()
    grids = 2 * num_descs * B  # target + reference volume per description
    # forward + backward (data + weight gradients) = 3x the forward convolution FLOPs
    tf = world * grids * 3 * UNET_GF_PER_GRID[C] / 1e3 / (ms / 1e3)
    return {"metric": "VOOL train steps/s (128^3, 16 ch, 16 descriptions/GPU)", "value": world * 1e3 / ms, "unit": "steps/s (sum over GPUs)",
            "precision": "fp32-accurate (3-term fp16 hi/lo convolutions forward and data-gradient)" if precise else
                         "single fp16 operands, fp32 accumulation (the counterpart of the reference's --use_amp)",
            "ms_per_step": ms, "voxel_grids_trained_per_s": world * grids / (ms / 1e3), "descs_per_gpu": num_descs,
            "unet_kernel_launches_per_step": launches_per_step,
            "loss_trajectory": loss_values, "unet_algorithmic_tflops": tf, "tensor_frac": tf / world / pk["tflops"],
            "peak_memory_gb": peak_gb,
            "note": ("gradient all-reduce: one NCCL bucket per UNet level, started from the backward pass as the level's weight "
                     "gradients are queued (train.GradientBuckets); the few parameters outside the UNet in one flat call after "
                     "backward") if world > 1 else "single GPU: no collective"}


if __name__ == "__main__":
    main()
