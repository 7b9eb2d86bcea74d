"""CPU suite for the training-side boundary: flag table, checkpoint key conventions, the synthetic batch
contract, the learning-rate schedule and `utils.loop`'s cross-rank stat averaging (gloo, world 2) — against
tests/golden/train_golden.json, which oracle/gen_golden_train.py produced by running the unmodified reference."""
import json
import os
import socket

import numpy as np
import pandas as pd
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "train_golden.json")))


def test_config_parser_matches_reference_flags_and_defaults():
    from semabs_b200 import utils

    ours = vars(utils.config_parser().parse_args(["--file_path", "x"]))
    ref = GOLD["config_defaults"]
    assert set(ours) == set(ref), set(ours) ^ set(ref)
    for k, v in ref.items():
        assert ours[k] == v, (k, ours[k], v)


def test_golden_frame_columns_are_what_get_losses_builds():
    from semabs_b200 import train_ovssc, train_vool  # noqa: F401  (import check: both modules expose the reference names)

    for mod in (train_ovssc, train_vool):
        assert callable(mod.get_losses) and callable(mod.get_detailed_stats) and "semantic_abstraction" in mod.approach
    m = ("precision", "recall", "false_negative", "false_positive", "iou")
    assert GOLD["ovssc_bal0"]["frame"]["columns"] == ["scene_id", "label"] + [f"point_{k}" for k in m] + \
        [f"voxel32x32x32_{k}" for k in m] + ["cutoff"]
    assert GOLD["vool_bal0"]["frame"]["columns"] == ["scene_id", "target_obj_name", "reference_obj_name", "spatial_relation_name"] + \
        [f"point_{k}" for k in m] + [f"voxel32x32x32_{k}" for k in m] + ["cutoff"]


def test_module_prefix_rule_and_checkpoint_layout():
    from semabs_b200 import utils

    sd = {"module.steps": 1, "module.vol_feature_extractor.final_conv.bias": 2, "steps2": 3,
          "module.completion_net.module.x": 4}
    out = utils.strip_module_prefix(sd)
    # reference rule (utils.py:282-287): drop everything up to the FIRST "module."; un-prefixed keys pass through
    assert out == {"steps": 1, "vol_feature_extractor.final_conv.bias": 2, "steps2": 3, "completion_net.module.x": 4}

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(2, 2)
            self.register_buffer("steps", torch.zeros(1))

    net = Net()
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    ck = utils.checkpoint_state(net, opt, 7)
    assert set(ck) == {"net", "optimizer", "epochs"} and ck["epochs"] == 7
    assert set(ck["net"]) == {"lin.weight", "lin.bias", "steps"}  # no process group -> no DDP prefix


def test_synthetic_datasets_collate_to_the_batch_contract():
    from torch.utils.data import DataLoader

    from semabs_b200 import utils

    ds = utils.SyntheticOVSSCDataset(length=3, num_input_pts=50, num_output_pts=70, num_patches=4)
    b = next(iter(DataLoader(ds, batch_size=2)))
    assert b["input_xyz_pts"].shape == (2, 50, 3) and b["input_feature_pts"].shape == (2, 4, 50, 1)
    assert b["output_xyz_pts"].shape == (2, 4, 70, 3) and b["output_label_pts"].shape == (2, 4, 70)
    assert b["out_of_frustum_pts_mask"].dtype == torch.bool and b["tsdf_vol"].shape == (2, 1)
    assert np.array(b["patch_labels"]).T.shape == (2, 4) and np.array(b["patch_labels"]).T[0, 3] == ""
    dv = utils.SyntheticVOOLDataset(length=3, num_input_pts=50, num_output_pts=70, num_descs=5)
    b = next(iter(DataLoader(dv, batch_size=2)))
    assert b["input_target_saliency_pts"].shape == (2, 5, 50, 1) and b["output_xyz_pts"].shape == (2, 5, 70, 3)
    assert np.array(b["spatial_relation_name"]).T.shape == (2, 5)
    again = next(iter(DataLoader(dv, batch_size=2)))
    assert torch.equal(again["output_xyz_pts"], b["output_xyz_pts"])  # seeded


def test_lr_schedule_matches_reference_get_net():
    """utils.get_net's scheduler (HF get_scheduler on the LAMB optimiser, utils.py:265-273): same learning rates."""
    from transformers import get_scheduler

    from semabs_b200.train import Lamb

    p = torch.nn.Parameter(torch.zeros(1))
    opt = Lamb([p], lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-5, adam=False)
    sched = get_scheduler("cosine_with_restarts", optimizer=opt, num_warmup_steps=4, num_training_steps=20)
    lrs = []
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # scheduler stepped without optimizer.step(): only the lr sequence is under test
        for _ in range(20):
            sched.step()
            lrs.append(opt.param_groups[0]["lr"])
    assert np.allclose(lrs, GOLD["lr_schedule_cosine_with_restarts_w4_t20"], rtol=1e-12, atol=0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("steps", torch.zeros(1))
        self.device = "cpu"


def _loop_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from semabs_b200 import utils

    def get_losses_fn(net, batch, **kw):  # evaluation branch: stats + a per-item frame, like the real get_losses
        v = float(batch["x"].sum())
        return {"loss": torch.tensor(v), "point_iou": v / 10}, pd.DataFrame({"scene_id": [f"r{rank}_{int(v)}"], "point_iou": [v / 10]})

    seen = {}

    class Logger:
        def add_scalar(self, k, v, step):
            seen.setdefault(k, []).append(v)

    loader = [{"x": torch.tensor([1.0 + rank])}, {"x": torch.tensor([3.0 + rank])}]
    df = utils.loop(net=_Net(), loader=loader, pbar=None, get_losses_fn=get_losses_fn, logger=Logger(), optimizer=None,
                    device=torch.device("cpu"))
    out.put((rank, sorted(df["scene_id"].tolist()), seen))
    dist.destroy_process_group()


def test_loop_averages_stats_and_gathers_frames_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_loop_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted((q.get(timeout=180) for _ in procs), key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    for rank, ids, seen in res:
        # every rank ends up with every rank's rows (all_gather_object, utils.py:433-435)
        assert ids == ["r0_1", "r0_3", "r1_2", "r1_4"]
    # rank 0 logs the epoch means of the RANK-AVERAGED stats: loss (1+2)/2 and (3+4)/2 -> mean 2.5
    assert abs(res[0][2]["loss_mean"][0] - 2.5) < 1e-12 and abs(res[0][2]["point_iou_mean"][0] - 0.25) < 1e-12
