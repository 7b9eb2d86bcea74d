"""Stand-alone stress of semabs_conv3d_halo's CTA-pair path at full size (debug aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semabs_b200 import ops

dev = "cuda"
def run(N, D, H, Ci, Co, precise, iters=6, stats=True, pair=True):
    ops.set_halo_pair(pair)
    s = 2 if precise else 1
    g = torch.Generator(device=dev).manual_seed(1)
    S = D * H * 128
    xn = (torch.randn(N * s * Ci * S, device=dev, generator=g) * 0.5).half()
    w = torch.randn(Co, Ci, 3, 3, 3, device=dev, generator=g) / (27 * Ci) ** 0.5
    wimg = ops.pack_halo_weights(w, s)
    out = torch.empty(N, S, Co, device=dev)
    st = torch.zeros(N, 8, 2, device=dev, dtype=torch.float64) if stats else None
    res = []
    for it in range(iters):
        t0 = time.time()
        ops.conv3d_halo(xn, wimg, N=N, D=D, H=H, W=128, C_in=Ci, C_out=Co, a_splits=s, w_splits=s, precise=precise, relu=True,
                        out32=out, stats=st, groups=8 if stats else 0)
        torch.cuda.synchronize()
        import ctypes as C
        from semabs_b200._lib import lib
        buf = (C.c_int32 * 512)()
        nrec = lib().semabs_debug_halo_pair_dump(buf)
        if nrec:
            tags = {1: "plane_empty", 2: "plane_full(start)", 3: "tmem_empty", 4: "plane_full", 5: "tmem_full"}
            print(f"  !! {nrec} time-outs:", flush=True)
            for k in range(min(nrec, 24)):
                r = buf[8 * k : 8 * k + 8]
                print(f"     cta {r[0]} rank {r[1]} warp {r[2]} waits {tags.get(r[3], r[3])}[{r[4]}] parity {r[5]} extra {r[6]}", flush=True)
        res.append(out.double().sum().item())
        print(f"  N={N} D={D} H={H} Ci={Ci} precise={precise} pair={pair} iter {it}: ok {time.time()-t0:.4f}s checksum {res[-1]:.6e}", flush=True)
    return res

cases = [(4, 128, 128, 16, 32, True), (1, 128, 128, 32, 32, True), (4, 128, 128, 32, 32, True), (4, 128, 128, 32, 32, False), (4, 128, 128, 16, 32, True),
         (4, 128, 128, 16, 32, False), (2, 128, 128, 32, 32, True), (4, 128, 128, 32, 32, True)]
sel = sys.argv[1:] or None
for c in cases:
    a = run(*c, pair=True)
    b = run(*c, pair=False, iters=1)
    print("  match single-CTA:", abs(a[0] - b[0]) <= 1e-6 * abs(b[0]), flush=True)
