"""Debug helper: count ReLU-branch disagreements (ours vs oracle) in the training-mode forward of the VOOL test."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import unet_oracle
from semabs_b200 import unet3d_bwd
from semabs_b200.net import SemAbsVOOL
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_train_gpu import _semabs_args, _points, BOUNDS

dev = "cuda"
torch.manual_seed(31)
v = SemAbsVOOL(pointing_method="cosine_sim", pointing_dim=64, decoder_concat_xyz_pts=True, **_semabs_args()).to(dev)
B, D, n_in, n_out = 1, 3, 2000, 4000
xyz, _, oxyz = _points(32, B, D, n_in, n_out)
g = torch.Generator().manual_seed(33)
tgt, refsal = torch.randn(B, D, n_in, 1, generator=g), torch.randn(B, D, n_in, 1, generator=g)
rel = [["behind"], ["on the left of"], ["behind"]]
sd = {k: t.detach().cpu().clone() for k, t in v.state_dict().items()}
relus = []
orig_relu = F.relu
def relu(x, *a, **k):
    y = orig_relu(x, *a, **k); relus.append((x.detach(), y.detach())); return y
F.relu = relu
with torch.no_grad():
    out_ref = unet_oracle.semabsvool_forward(sd, xyz, tgt, refsal, oxyz, rel, BOUNDS, (16, 16, 16), concat_xyz=True)
F.relu = orig_relu
tapes = []
orig_new = unet3d_bwd.UNetBackward.new_tape
def new_tape(self):
    t = orig_new(self); tapes.append(t); return t
unet3d_bwd.UNetBackward.new_tape = new_tape
out = v(output_xyz_pts=oxyz.to(dev), spatial_relation_name=rel, input_xyz_pts=xyz.to(dev), input_target_saliency_pts=tgt.to(dev),
        input_reference_saliency_pts=refsal.to(dev), tsdf_vol=torch.ones(B, 1, device=dev))
print("training-mode forward logits err", ((out.detach().cpu() - out_ref).abs().max() / out_ref.abs().max()).item())
names = [f"{p}.{k}" for p in ("enc0", "enc1", "enc2", "dec0", "dec1") for k in ("o1", "o2", "out")]
assert len(relus) == 2 * len(names), len(relus)
total = 0
for pi, tape in enumerate(tapes):
    for ni, name in enumerate(names):
        pre, ref = relus[pi * len(names) + ni]
        blk, key = name.split(".")
        rec = tape.blocks[blk]
        ours = rec[key].view(3, *rec["dims"], -1).permute(0, 4, 1, 2, 3).cpu()
        err = ((ours - ref).abs().max() / ref.abs().max()).item()
        flips = ((ours > 0) != (ref > 0))
        nf = int(flips.sum())
        total += nf
        if nf or err > 1e-4:
            print(f"pass {pi} {name}: err {err:.2e}, branch flips {nf} of {flips.numel()}, |pre| at flips {pre[flips].abs().tolist()[:5]} (rms {pre.pow(2).mean().sqrt():.3f})")
print("total ReLU branch flips:", total)
