"""Host-side mirror of the reference's `CLIP.clip` public surface for the relevancy path: `saliency_configs`,
`ClipWrapper` (singleton, same classmethods and keyword arguments; reference CLIP/clip/__init__.py:19-282) and
`ClipGradcam` (clip_gradcam.py:30-142).  Host work is limited to what the reference also does on the host
(PIL crop/resize/normalise, tokenisation, tile enumeration); every tensor op runs in libsemabs_b200.so.
"""
from __future__ import annotations

import itertools
from typing import List, Sequence

import numpy as np
import torch
from PIL import Image

from .. import ops
from .engine import ClipEngine
from .model import load_state_dict, pack_clip_weights
from .tokenizer import tokenize

saliency_configs = {
    # same pyramid / TTA specs as the reference (CLIP/clip/__init__.py:19-41)
    "ours": lambda img_dim: {
        "distractor_labels": {},
        "horizontal_flipping": True,
        "augmentations": 5,
        "imagenet_prompt_ensemble": False,
        "positive_attn_only": True,
        "cropping_augmentations": [
            {"tile_size": img_dim, "stride": img_dim // 4},
            {"tile_size": int(img_dim * 2 / 3), "stride": int(img_dim * 2 / 3) // 4},
            {"tile_size": img_dim // 2, "stride": (img_dim // 2) // 4},
            {"tile_size": img_dim // 4, "stride": (img_dim // 4) // 4},
        ],
    },
    "chefer_et_al": lambda img_dim: {
        "distractor_labels": {},
        "horizontal_flipping": False,
        "augmentations": 0,
        "imagenet_prompt_ensemble": False,
        "positive_attn_only": True,
        "cropping_augmentations": [{"tile_size": img_dim, "stride": img_dim // 4}],
    },
}

_MEAN = torch.tensor((0.48145466, 0.4578275, 0.40821073)).view(1, 3, 1, 1)
_STD = torch.tensor((0.26862954, 0.26130258, 0.27577711)).view(1, 3, 1, 1)


_POOL = None


def _pool():
    global _POOL
    if _POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor

        _POOL = ThreadPoolExecutor(max_workers=max(1, min(32, (os.cpu_count() or 2))))
    return _POOL


def preprocess_tiles(tiles_u8: Sequence[np.ndarray], n_px: int) -> torch.Tensor:
    """Host tile preprocessing, fanned out over a thread pool (PIL's resize releases the GIL); the per-tile
    arithmetic is `_preprocess_tiles_serial`'s."""
    n = len(tiles_u8)
    if n < 32:
        return _preprocess_tiles_serial(tiles_u8, n_px)
    step = max(8, (n + 63) // 64)
    chunks = [tiles_u8[i : i + step] for i in range(0, n, step)]
    return torch.cat(list(_pool().map(lambda c: _preprocess_tiles_serial(c, n_px), chunks)), dim=0)


def _preprocess_tiles_into(out: torch.Tensor, tiles_u8: Sequence[np.ndarray], n_px: int) -> torch.Tensor:
    """The reference's `_transform` (clip_explainability.py:98-108) on a list of uint8 crops, written into `out`
    [len(tiles), 3, n_px, n_px]: PIL bicubic resize of the shorter side to 224 (hard-coded there), centre crop/pad to
    n_px, /255, normalise. Host side, like the reference (it is the reference's declared bottleneck, __init__.py:275;
    a device version is a 'next' row)."""
    for i, t in enumerate(tiles_u8):
        img = Image.fromarray(t).convert("RGB")
        w, h = img.size
        nw, nh = (224, int(224 * h / w)) if w <= h else (int(224 * w / h), 224)
        arr = torch.from_numpy(np.array(img.resize((nw, nh), Image.BICUBIC), dtype=np.uint8)).permute(2, 0, 1)
        if (nh, nw) != (n_px, n_px):
            if nh < n_px or nw < n_px:
                pl, pt = max((n_px - nw) // 2, 0), max((n_px - nh) // 2, 0)
                pr, pb = max((n_px - nw + 1) // 2, 0), max((n_px - nh + 1) // 2, 0)
                arr = torch.nn.functional.pad(arr, (pl, pr, pt, pb))
                nh, nw = arr.shape[1:]
            top, left = int(round((nh - n_px) / 2.0)), int(round((nw - n_px) / 2.0))
            arr = arr[:, top : top + n_px, left : left + n_px]
        # (x/255 - mean)/std with the same operation order as ToTensor + Normalize
        out[i] = (arr.float().div(255) - _MEAN[0]) / _STD[0]
    return out


def _preprocess_tiles_serial(tiles_u8: Sequence[np.ndarray], n_px: int) -> torch.Tensor:
    return _preprocess_tiles_into(torch.empty(len(tiles_u8), 3, n_px, n_px), tiles_u8, n_px)


_PRECISION_BITS = 32 - 8 - 2  # Pillow: src/libImaging/Resample.c


def _bicubic_filter(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pillow_bicubic_coeffs(in_size: int, out_size: int):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for Image.resize(..., BICUBIC) along one axis (what
    `Resize(224, interpolation=BICUBIC)` of the reference's `_transform` runs, clip_explainability.py:98-108):
    -> (coef int64 [out_size, ksize] 22-bit fixed point, bounds int64 [out_size, 2] = (first input index, window length)).
    Same double-precision expressions in the same order as the C source; pinned bit-exactly against the installed Pillow in
    tests/test_abi.py::test_pillow_resize_tables."""
    import math

    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    coef = np.zeros((out_size, ksize), np.int64)
    bounds = np.zeros((out_size, 2), np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            coef[xx, x] = int(k * (1 << _PRECISION_BITS) - 0.5) if k < 0 else int(k * (1 << _PRECISION_BITS) + 0.5)
        bounds[xx] = (xmin, xmax)
    return coef, bounds


_COEF_KMAX = 32  # window lengths up to 32 taps: tiles up to ~1600 px per side


class _PinnedRing:
    """A few pinned host batches [tile_batch, 3, R, R] that worker threads fill slice-wise and that are copied to the
    device asynchronously; a buffer is reused only after its previous H2D copy has completed (CUDA event)."""

    DEPTH = 4

    def __init__(self, shape):
        self.shape = tuple(shape)
        self.bufs = [torch.empty(self.shape).pin_memory() for _ in range(self.DEPTH)]
        self.events = [None] * self.DEPTH
        self.i = 0

    def acquire(self):
        k = self.i
        self.i = (k + 1) % self.DEPTH
        if self.events[k] is not None:
            self.events[k].synchronize()
            self.events[k] = None
        return k, self.bufs[k]

    def copied(self, k):
        self.events[k] = torch.cuda.Event()
        self.events[k].record()


class ClipGradcam:
    """Drop-in for the reference class of the same name: `ClipGradcam(...)(x, o)` -> relevance [P,B,g,g]."""

    def __init__(self, clip_model_name: str, classes: List[str], templates: List[str], device, num_layers=10,
                 positive_attn_only=False, fwd_splits=2, bwd_splits=1, seed=None, allow_synthetic=None, **load_kwargs):
        sd = load_state_dict(clip_model_name, load_kwargs.get("download_root"), seed=seed, allow_synthetic=allow_synthetic)
        self.synthetic = "__synthetic__" in sd
        sd.pop("__synthetic__", None)
        self.clip_model_name = clip_model_name
        self.device = torch.device(device)
        self.weights = pack_clip_weights(clip_model_name, sd, self.device, first_rollout_block=num_layers + 1)
        self.engine = ClipEngine(self.weights, self.device, num_layers=num_layers, fwd_splits=fwd_splits,
                                 bwd_splits=bwd_splits)
        self.n_px = self.weights.input_resolution
        self.preprocess = lambda pil_img: preprocess_tiles([np.array(pil_img)], self.n_px)[0]
        self.templates = templates
        self.num_layers = num_layers
        self.positive_attn_only = positive_attn_only
        self.target_classes = None
        self.class_to_language_feature = {}
        self.set_classes(classes)

    def set_classes(self, classes):
        """zeroshot_classifier + dict keyed by label (duplicates collapse, as in clip_gradcam.py:134-142)."""
        self.target_classes = classes
        texts = list(itertools.chain(*[[t.format(c) for t in self.templates] for c in classes]))
        W = self.engine.zeroshot_weights(tokenize(texts), len(classes), len(self.templates))
        self.class_to_language_feature = {c: W[:, [i]] for i, c in enumerate(classes)}

    def __call__(self, x: torch.Tensor, o: Sequence[str]):
        W = torch.cat([self.class_to_language_feature[p] for p in o], dim=1).contiguous()
        x = x.to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
        return self.engine.relevancy(x, W, positive_attn_only=self.positive_attn_only)

    forward = __call__


class ClipWrapper:
    # SINGLETON WRAPPER (class-level state like the reference, __init__.py:44-51: one instance per process)
    clip_gradcam = None
    device = None
    jittering_transforms = None
    engine_kwargs: dict = {}

    def __init__(self, clip_model_type, device, **kwargs):
        import torchvision

        ClipWrapper.device = torch.device(device)
        ClipWrapper.jittering_transforms = torchvision.transforms.ColorJitter(
            brightness=0.6, contrast=0.6, saturation=0.6, hue=0.1
        )
        ClipWrapper.clip_gradcam = ClipGradcam(
            clip_model_name=clip_model_type, classes=[""], templates=["{}"], device=ClipWrapper.device, **kwargs
        )

    @classmethod
    def check_initialized(cls, clip_model_type="ViT-B/32", **kwargs):
        if cls.clip_gradcam is None:
            if not torch.cuda.is_available():
                raise RuntimeError("semabs_b200.ClipWrapper needs a CUDA device: there is no CPU path")
            ClipWrapper(clip_model_type=clip_model_type, device="cuda", **kwargs)

    @classmethod
    def reset(cls):
        cls.clip_gradcam = None

    @classmethod
    def get_clip_text_feature(cls, string):
        cls.check_initialized()
        return cls.clip_gradcam.engine.encode_text(tokenize(string, context_length=77)).squeeze().cpu().numpy()

    @classmethod
    def get_clip_saliency(cls, img, text_labels, prompts, distractor_labels=set(), use_lavt=False, **kwargs):
        """Same contract as the reference (__init__.py:103-133): returns (maps [P,H,W] fp32 on CPU,
        text features [P,E] on CPU)."""
        cls.check_initialized()
        if use_lavt:
            raise NotImplementedError("LAVT localisation is outside the hot path (SURVEY.md §8)")
        cls.clip_gradcam.templates = prompts
        cls.clip_gradcam.set_classes(list(text_labels))
        text_label_features = torch.stack(list(cls.clip_gradcam.class_to_language_feature.values()), dim=0)
        text_label_features = text_label_features.squeeze(dim=-1).cpu()
        text_maps = cls.get_clip_saliency_convolve(img=img, text_labels=text_labels, **kwargs)
        if len(distractor_labels) > 0:
            distractor_labels = set(distractor_labels) - set(text_labels)
            cls.clip_gradcam.set_classes(list(distractor_labels))
            distractor_maps = cls.get_clip_saliency_convolve(img=img, text_labels=list(distractor_labels), **kwargs)
            text_maps -= distractor_maps.mean(dim=0)
        return text_maps.cpu(), text_label_features.squeeze(dim=-1)

    @classmethod
    def get_clip_saliency_device(cls, tile_imgs, tile_desc, size_order, text_labels, H, W, horizontal_flipping=False,
                                 positive_attn_only=False, tile_batch_size=32, prompt_batch_size=32):
        """Device half of get_clip_saliency_convolve: preprocessed tiles -> maps [P,H,W] on the device.
        `tile_imgs` is a tensor [n,3,R,R] (host or device) or an iterable of such batches (e.g. produced by worker
        threads while earlier batches are already on the GPU)."""
        gc = cls.clip_gradcam
        gc.positive_attn_only = positive_attn_only
        dev = cls.device
        labels = list(text_labels)
        if torch.is_tensor(tile_imgs):
            tile_imgs = [tile_imgs[t : t + tile_batch_size] for t in range(0, len(tile_imgs), tile_batch_size)]

        def relevancy(tiles):  # [b,3,R,R] on device -> [P, b, g, g]
            return torch.cat([gc(x=tiles, o=labels[p : p + prompt_batch_size]).clone()
                              for p in range(0, len(labels), prompt_batch_size)], dim=0)  # fmt: skip

        rel_parts, kept = [], []
        for batch in tile_imgs:
            batch = cls._to_device(batch)
            rel_parts.append(relevancy(batch))
            if horizontal_flipping:
                kept.append(batch)
        rel = torch.cat(rel_parts, dim=1).contiguous()
        if horizontal_flipping:
            # flipping the tile is pure data movement (__init__.py:170-173)
            flipped = torch.cat([relevancy(b.flip(-1).contiguous()) for b in kept], dim=1).contiguous()
            rel = ops.flip_average(rel, flipped)
        out = torch.empty(len(labels), H, W, device=dev)
        desc = torch.as_tensor(np.ascontiguousarray(tile_desc), dtype=torch.int32).to(dev)
        order = torch.as_tensor(list(size_order), dtype=torch.int32).to(dev)
        return ops.tile_assemble(rel, desc, order, H, W, out)

    _staging = None
    _last_images = None

    @classmethod
    def _to_device(cls, batch):
        """Host batch -> device through a small ring of pinned staging buffers (async H2D on the current stream)."""
        if batch.is_cuda:
            return batch
        if batch.is_pinned():
            return batch.to(cls.device, non_blocking=True)
        if cls._staging is None or cls._staging["buf"][0].shape[1:] != batch.shape[1:] or cls._staging["buf"][0].shape[0] < batch.shape[0]:
            cls._staging = {"buf": [torch.empty((max(32, batch.shape[0]),) + tuple(batch.shape[1:])).pin_memory() for _ in range(3)],
                            "ev": [None] * 3, "i": 0}
        st = cls._staging
        i = st["i"]
        st["i"] = (i + 1) % 3
        if st["ev"][i] is not None:
            st["ev"][i].synchronize()
        pinned = st["buf"][i][: batch.shape[0]]
        pinned.copy_(batch)
        out = pinned.to(cls.device, non_blocking=True)
        st["ev"][i] = torch.cuda.Event()
        st["ev"][i].record()
        return out

    device_preprocessing = True  # crop / Pillow-exact resize / normalise on the GPU (semabs_tile_preprocess)
    _coef_cache: dict = {}

    @classmethod
    def _device_preprocessed_batches(cls, tile_desc, n_px, tile_batch_size):
        """Tile preprocessing on the device: the uint8 image copies go up once (0.3 MB each instead of 0.6 MB per TILE of
        fp32), every tile batch is one launch of semabs_tile_preprocess — bit-identical to the PIL host path (tests).
        Returns None when the kernel's preconditions do not hold (non-224 input resolution, windows above 32 taps):
        the host path, which is also what the reference does, takes over."""
        images = cls._last_images
        sizes = sorted({int(s) for s in tile_desc[:, 2]})
        if n_px != 224 or not sizes or any(2 * int(np.ceil(2.0 * max(s / 224, 1.0))) + 1 > _COEF_KMAX for s in sizes):
            return None
        dev = cls.device
        key = tuple(sizes)
        if key not in cls._coef_cache:
            coef = np.zeros((len(sizes), 224, _COEF_KMAX), np.int32)
            bounds = np.zeros((len(sizes), 224, 2), np.int32)
            for i, sz in enumerate(sizes):
                c, b = pillow_bicubic_coeffs(sz, 224)
                coef[i, :, : c.shape[1]] = c
                bounds[i] = b
            cls._coef_cache[key] = (torch.from_numpy(coef).to(dev), torch.from_numpy(bounds).to(dev))
        coef_d, bounds_d = cls._coef_cache[key]
        n = len(tile_desc)
        n_images = len(images) + len(cls._jitter_params)
        per_img = n // n_images
        table = np.zeros((n, 5), np.int32)
        table[:, 0] = np.arange(n) // per_img
        table[:, 1:4] = tile_desc
        table[:, 4] = np.searchsorted(np.asarray(sizes), tile_desc[:, 2])
        imgs_d = torch.from_numpy(np.ascontiguousarray(np.stack(images))).pin_memory().to(dev, non_blocking=True)
        if cls._jitter_params:  # the augmentation copies, made here on the device from the uploaded original
            imgs_d = torch.stack([imgs_d[0]] + [ops.color_jitter(imgs_d[0], *prm) for prm in cls._jitter_params])
        table_d = torch.from_numpy(table).pin_memory().to(dev, non_blocking=True)
        mean, std = [float(v) for v in _MEAN.flatten()], [float(v) for v in _STD.flatten()]

        def gen():
            for i in range(0, n, tile_batch_size):
                cnt = min(tile_batch_size, n - i)
                out = torch.empty(cnt, 3, 224, 224, device=dev)
                yield ops.tile_preprocess(imgs_d, table_d[i : i + cnt], coef_d, bounds_d, out, mean, std)

        return gen()

    _ring = None
    SUB_TILES = 12  # tiles per worker task: a 95-tile batch is preprocessed by 8 threads at once, not by one

    @classmethod
    def _device_batches(cls, crops, n_px, tile_batch_size):
        """Generator of preprocessed tile batches already on their way to the device.  Every batch is cut into
        SUB_TILES-sized worker tasks that fill slices of one pinned buffer, so the FIRST batch of an image is ready after
        ~SUB_TILES tile-times instead of tile_batch_size (that latency was 140 ms of idle GPU per image); up to
        DEPTH-1 batches are prepared ahead of the one the GPU is working on."""
        from collections import deque

        n = len(crops)
        shape = (tile_batch_size, 3, n_px, n_px)
        if cls._ring is None or cls._ring.shape != shape:
            cls._ring = _PinnedRing(shape)
        ring, pool = cls._ring, _pool()
        starts = deque(range(0, n, tile_batch_size))
        pending = deque()

        def submit():
            i = starts.popleft()
            cnt = min(tile_batch_size, n - i)
            k, buf = ring.acquire()
            futs = [pool.submit(_preprocess_tiles_into, buf[j : min(j + cls.SUB_TILES, cnt)], crops[i + j : i + min(j + cls.SUB_TILES, cnt)], n_px)
                    for j in range(0, cnt, cls.SUB_TILES)]
            pending.append((k, buf, cnt, futs))

        while starts and len(pending) < ring.DEPTH - 1:
            submit()
        while pending:
            k, buf, cnt, futs = pending.popleft()
            for f in futs:
                f.result()
            d = buf[:cnt].to(cls.device, non_blocking=True)
            ring.copied(k)
            if starts:
                submit()
            yield d

    @classmethod
    def get_clip_saliency_convolve(cls, text_labels, horizontal_flipping=False, positive_attn_only: bool = False,
                                   tile_batch_size=32, prompt_batch_size=32, tile_interpolate_batch_size=32, **kwargs):
        """Reference: CLIP/clip/__init__.py:135-236. Host tile preprocessing runs in worker threads and overlaps with the
        GPU work of the batches already submitted (see _device_batches)."""
        tile_desc, crops, size_order = cls.enumerate_crops(**kwargs)
        n_px = cls.clip_gradcam.n_px
        H, W = kwargs["img"].shape[:2]
        batches = cls._device_preprocessed_batches(tile_desc, n_px, tile_batch_size) if cls.device_preprocessing else None
        if batches is None:
            if cls._jitter_params:  # host path: make the SAME jitter copies with PIL, like the reference
                tile_desc, crops, size_order = cls.enumerate_crops(host_jitter=True, jitter_params=cls._all_jitter_params, **kwargs)
            batches = cls._device_batches(crops, n_px, tile_batch_size)
        out = cls.get_clip_saliency_device(batches, tile_desc, size_order, text_labels, H, W, horizontal_flipping,
                                           positive_attn_only, tile_batch_size, prompt_batch_size)
        return out if kwargs.get("keep_on_device", False) else out.cpu()

    device_jitter = True  # ColorJitter copies of the "ours" TTA made on the GPU (semabs_color_jitter_op), not by PIL on the host
    _jitter_params: list = []

    @staticmethod
    def _jitter_on_host(img_pil, params):
        """what ColorJitter.forward does to a PIL image for drawn parameters (the reference's path, __init__.py:246-247)"""
        import torchvision.transforms.functional as TF

        fn_idx, b, c, s, h = params
        for fn_id in fn_idx:
            if fn_id == 0 and b is not None:
                img_pil = TF.adjust_brightness(img_pil, b)
            elif fn_id == 1 and c is not None:
                img_pil = TF.adjust_contrast(img_pil, c)
            elif fn_id == 2 and s is not None:
                img_pil = TF.adjust_saturation(img_pil, s)
            elif fn_id == 3 and h is not None:
                img_pil = TF.adjust_hue(img_pil, h)
        return img_pil

    @classmethod
    def enumerate_crops(cls, img, augmentations, cropping_augmentations, host_jitter=None, jitter_params=None, **kwargs):
        """Tile enumeration in the reference's order (__init__.py:238-282): image copies (original + ColorJitter
        draws) -> crop sizes -> column offset -> row offset. Returns (tile_desc int32 [n,3] = (row0,col0,size),
        list of uint8 crops (views), size order).  The jitter parameters are drawn exactly like ColorJitter.forward draws
        them (same torch RNG consumption as the reference); with `device_jitter` the copies themselves are produced on the
        GPU later (`_device_preprocessed_batches`) and the crop list only covers the original image."""
        assert type(img) == np.ndarray
        cls.check_initialized()
        jt = cls.jittering_transforms
        if host_jitter is None:
            host_jitter = not (cls.device_preprocessing and cls.device_jitter)
        if jitter_params is None:
            jitter_params = [jt.get_params(jt.brightness, jt.contrast, jt.saturation, jt.hue) for _ in range(augmentations)]
        img_pil = Image.fromarray(img)
        images = [np.array(img_pil)]
        if host_jitter:
            images += [np.array(cls._jitter_on_host(img_pil, prm)) for prm in jitter_params]
        desc, crops = [], []
        for k in range(1 + augmentations):
            im = images[k] if k < len(images) else None
            H_, W_ = images[0].shape[:2]
            for aug in cropping_augmentations:
                ts, st = aug["tile_size"], aug["stride"]
                for y in np.arange(0, W_ - ts + 1, st):
                    if y >= H_:
                        continue
                    for x in np.arange(0, H_ - ts + 1, st):
                        if x >= W_:
                            continue
                        desc.append((int(x), int(y), int(ts)))
                        if im is not None:
                            crops.append(im[x : x + ts, y : y + ts])
        size_order = list(dict.fromkeys(a["tile_size"] for a in cropping_augmentations))
        cls._last_images = images  # the device preprocessing path uploads these instead of the tiles
        cls._jitter_params = [] if host_jitter else list(jitter_params)
        cls._all_jitter_params = list(jitter_params)
        return np.array(desc, dtype=np.int32).reshape(-1, 3), crops, size_order

    @classmethod
    def create_tiles(cls, img, augmentations, cropping_augmentations, **kwargs):
        """Eager variant of the reference's create_tiles: (tile_desc, preprocessed tiles [n,3,R,R] fp32 in pinned host
        memory, size order)."""
        desc, crops, size_order = cls.enumerate_crops(img, augmentations, cropping_augmentations, host_jitter=True)
        tiles = preprocess_tiles(crops, cls.clip_gradcam.n_px)
        if torch.cuda.is_available():
            tiles = tiles.pin_memory()
        return desc, tiles, size_order
