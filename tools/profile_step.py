"""One relevancy batch (32 tiles x 16 labels, ViT-L/14) and one UNet forward (128^3 x 32ch, batch 4) for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
what = sys.argv[1] if len(sys.argv) > 1 else "vit"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if what == "vit":
    from semabs_b200.clip.engine import ClipEngine
    from semabs_b200.clip.model import pack_clip_weights, synthetic_clip_state_dict
    sd = synthetic_clip_state_dict("ViT-L/14", seed=0)
    eng = ClipEngine(pack_clip_weights("ViT-L/14", sd, "cuda"), "cuda")
    g = torch.Generator().manual_seed(0)
    NB = int(os.environ.get("VIT_B", "32"))
    tiles = torch.randn(NB, 3, 224, 224, generator=g).cuda()
    W = torch.randn(768, 16, generator=g); W = (W / W.norm(dim=0, keepdim=True)).contiguous().cuda()
    for _ in range(1 + reps):
        eng.relevancy(tiles, W)
    torch.cuda.synchronize()
else:
    from semabs_b200.unet3d import ResidualUNet3D
    C = int(os.environ.get("UNET_C", "32"))
    torch.manual_seed(0)
    m = ResidualUNet3D(in_channels=C, out_channels=C, f_maps=C, num_groups=8, num_levels=6, precise=os.environ.get("UNET_FAST", "0") != "1").cuda()
    x = torch.randn(4, C, 128, 128, 128, device="cuda")
    torch.set_grad_enabled(False)  # inference path (with grad enabled the module records a tape for backward)
    if os.environ.get("STEADY", "0") == "1":
        # steady state only (run under `ncu --profile-from-start off`): the input draw, the weight packs and the workspace
        # allocation of the first forwards stay outside the profiled window
        for _ in range(2):
            m(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for _ in range(reps):
            m(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    else:
        for _ in range(1 + reps):
            m(x)
    torch.cuda.synchronize()
if os.environ.get("TIME", "0") == "1":
    import time
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        (eng.relevancy(tiles, W) if what == "vit" else m(x))
    torch.cuda.synchronize(); print(f"{what}: {(time.perf_counter()-t0)/n*1e3:.1f} ms per call", flush=True)
