"""GPU parity of the callers around the network: sliding-window inference (trainer.py:17-98) against the oracle's
restatement, and the Trainer step loop / checkpoint round trip (trainer.py:415-634)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import unet3d_b200  # noqa: E402
from unet3d_b200 import ops  # noqa: E402
from oracle import unet3d_oracle as O  # noqa: E402

DEV = "cuda"


class _ToyNet(torch.nn.Module):
    """A stand-in 'model' for the blend tests: one fp32 conv evaluated with torch (the network itself is covered
    by test_model_gpu.py); makes the window arithmetic checkable to fp32 accuracy."""

    def __init__(self, w, b):
        super().__init__()
        self.w, self.b = torch.nn.Parameter(w), torch.nn.Parameter(b)

    def forward(self, x):
        return torch.nn.functional.conv3d(x, self.w, self.b, padding=1)


def test_predict_per_patch_matches_reference_golden(golden_dir):
    """Same toy conv + volume as tests/golden/predict_toy.npz (made by the live reference's predict_per_patch):
    labels bit-exact incl. the uncovered border (label 0), probabilities to 1e-6 with NaN where uncovered."""
    torch.backends.cudnn.allow_tf32 = False
    z = np.load(os.path.join(golden_dir, "predict_toy.npz"))
    net = _ToyNet(torch.from_numpy(z["w"]), torch.from_numpy(z["b"])).to(DEV)
    lab = unet3d_b200.predict_per_patch(z["vol"], net, 3, (16, 24, 16), 2, verbose=False)
    assert lab.dtype == np.uint8 and lab.shape == z["labels"].shape
    # BIT-EXACT label map (north_star: "label maps bit-exact").  The toy conv runs in fp32 on cuDNN here and on the CPU in
    # the reference; the smallest class margin of the fixture's blended probabilities is 2.8e-6, ~30x the fp32 rounding
    # of a 27-tap conv + softmax, so no voxel is allowed to differ
    assert np.array_equal(lab, z["labels"]), int((lab != z["labels"]).sum())
    prob = unet3d_b200.predict_per_patch(z["vol"], net, 3, (16, 24, 16), 2, verbose=False, one_hot=True)
    assert np.array_equal(np.isnan(prob), np.isnan(z["probs"]))
    assert np.allclose(np.nan_to_num(prob), np.nan_to_num(z["probs"]), atol=2e-6)


@pytest.mark.parametrize("window", [None, "gaussian"])
@pytest.mark.parametrize("grid_mode", ["reference", "full_cover"])
def test_predict_per_patch_windows_and_grids(window, grid_mode):
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(3)
    vol = torch.randn(40, 30, 37, 1, generator=g).numpy()
    w, b = torch.randn(3, 1, 3, 3, 3, generator=g), torch.randn(3, generator=g)
    net = _ToyNet(w, b).to(DEV)
    patch = (16, 16, 16)
    got = unet3d_b200.predict_per_patch(vol, net, 3, patch, 2, verbose=False, one_hot=True, window=window, grid_mode=grid_mode)
    # oracle with the same grid / window
    fn = lambda t: torch.nn.functional.conv3d(t, w, b, padding=1)
    win = None if window is None else O.gaussian_window(patch)
    if grid_mode == "reference":
        want = O.predict_per_patch(vol, fn, 3, patch, 2, one_hot=True, window=win)
    else:
        padded = O.pad_to(vol, patch)
        res = torch.zeros(3, *padded.shape[:3]); wsum = torch.zeros(padded.shape[:3])
        xx = torch.from_numpy(np.ascontiguousarray(np.moveaxis(padded, -1, 0))[None])
        for (ox, oy, oz) in unet3d_b200.tile_origins(padded.shape[:3], patch, 2, "full_cover"):
            p = torch.softmax(fn(xx[:, :, ox:ox + 16, oy:oy + 16, oz:oz + 16]), 1)[0]
            wt = torch.ones(patch) if win is None else torch.from_numpy(win)
            res[:, ox:ox + 16, oy:oy + 16, oz:oz + 16] += p * wt
            wsum[ox:ox + 16, oy:oy + 16, oz:oz + 16] += wt
        want = O.crop_pad(np.moveaxis((res / wsum).numpy(), 0, -1), vol.shape[:3])
        assert not np.isnan(want).any()               # full_cover reaches every voxel
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.allclose(np.nan_to_num(got), np.nan_to_num(want), atol=3e-6)


def test_predict_per_patch_with_the_unet():
    torch.manual_seed(0)
    model = unet3d_b200.ResUnet3D(num_pool=2, num_features=8, out_channels=3)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.to(DEV)
    model.precision = "fp16"
    vol = torch.randn(40, 36, 24, 1, generator=torch.Generator().manual_seed(1)).numpy()
    got = unet3d_b200.predict_per_patch(vol, model, 3, (16, 16, 16), 2, verbose=False, one_hot=True, grid_mode="full_cover")
    padded = vol
    res = torch.zeros(3, *padded.shape[:3]); wsum = torch.zeros(padded.shape[:3])
    xx = torch.from_numpy(np.ascontiguousarray(np.moveaxis(padded, -1, 0))[None])
    with torch.no_grad():
        for (ox, oy, oz) in unet3d_b200.tile_origins(padded.shape[:3], (16, 16, 16), 2, "full_cover"):
            p = torch.softmax(O.resunet3d_forward(sd, xx[:, :, ox:ox + 16, oy:oy + 16, oz:oz + 16], 2, 8), 1)[0]
            res[:, ox:ox + 16, oy:oy + 16, oz:oz + 16] += p
            wsum[ox:ox + 16, oy:oy + 16, oz:oz + 16] += 1
    want = np.moveaxis((res / wsum).numpy(), 0, -1)
    assert np.abs(got - want).max() < 2e-2
    assert (got.argmax(-1) == want.argmax(-1)).mean() > 0.995


class _Cases(torch.utils.data.Dataset):
    def __init__(self, n):
        g = torch.Generator().manual_seed(11)
        self.x = torch.randn(n, 1, 16, 16, 16, generator=g)
        self.y = (self.x[:, 0] > 0.3).long() + (self.x[:, 0] > 1.0).long()

    def __len__(self):
        return len(self.x)

    def __getitem__(self, i):
        return {"image": self.x[i], "label": self.y[i]}


def test_trainer_fit_learns_and_checkpoints(tmp_path):
    torch.manual_seed(0)
    np.random.seed(0)
    model = unet3d_b200.ResUnet3D(num_pool=2, num_features=8, out_channels=3).to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    tr = unet3d_b200.Trainer(model, opt, unet3d_b200.HybirdLoss(), _Cases(8), batch_size=2,
                             dataloader_kwargs={"num_workers": 0, "pin_memory": True}, valid_split=0.25,
                             metrics={"dice": unet3d_b200.Dice()})
    save = str(tmp_path / "ck")
    tr.fit(num_epochs=3, save_dir=save, use_amp=True)
    assert model.precision == "fp16"
    first = tr.batch_loop(torch.utils.data.DataLoader(_Cases(8), batch_size=2), is_train=False)
    assert set(first) == {"loss", "dice"} and np.isfinite(first["loss"])
    assert os.path.exists(save + "-last.pt") and os.path.exists(save + "-best.pt")
    # learns: a few more epochs lower the training loss
    l0 = tr.batch_loop(torch.utils.data.DataLoader(_Cases(8), batch_size=2), is_train=True)["loss"]
    for _ in range(6):
        l1 = tr.batch_loop(torch.utils.data.DataLoader(_Cases(8), batch_size=2), is_train=True)["loss"]
    assert l1 < l0
    # checkpoint round trip (reference dictionary layout)
    ck = torch.load(save + "-last.pt", weights_only=False)
    assert {"model_state_dict", "optimizer_state_dict", "current_epoch", "train_indices", "valid_indices",
            "best_result"} <= set(ck)
    m2 = unet3d_b200.ResUnet3D(num_pool=2, num_features=8, out_channels=3).to(DEV)
    tr2 = unet3d_b200.Trainer(m2, torch.optim.Adam(m2.parameters(), lr=2e-3), unet3d_b200.HybirdLoss(), _Cases(8), batch_size=2)
    tr2.load_checkpoint(save + "-last.pt")
    assert tr2.current_epoch == 3
    for k, v in ck["model_state_dict"].items():
        assert torch.equal(m2.state_dict()[k].cpu(), v.cpu())
    ops.check_device_errors()


@pytest.mark.parametrize("fused", [True, False])
def test_forward_sees_optimizer_updates(fused):
    """Packed 16-bit weight tiles are cached between forwards; the cache must not survive an optimizer step.
    torch.optim.Adam(fused=True) updates parameters WITHOUT bumping their version counters, so the version alone
    cannot be the cache key (ops.PACK_EPOCH).  Checked against the oracle run on the updated state_dict."""
    torch.manual_seed(0)
    m = unet3d_b200.ResUnet3D(num_pool=2, num_features=8, out_channels=3).to(DEV).eval()     # eval: no dropout
    opt = torch.optim.Adam(m.parameters(), lr=1e-2, fused=fused)
    x = torch.randn(1, 1, 16, 16, 16, device=DEV)
    y = torch.randint(0, 3, (1, 16, 16, 16), device=DEV)
    sd0 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    for _ in range(2):
        opt.zero_grad()
        unet3d_b200.DiceLoss()(m(x), y).backward()
        opt.step()
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    ref = O.resunet3d_forward(sd, x.cpu(), 2, 8)
    stale = O.resunet3d_forward(sd0, x.cpu(), 2, 8)
    assert ((stale - ref).norm() / ref.norm()).item() > 0.2          # the update is large enough to be seen
    with torch.no_grad():
        out = m(x).cpu()
    assert ((out - ref).norm() / ref.norm()).item() < 3e-2
    out2 = m(x).detach().cpu()                                       # grad-enabled forward after the step
    assert ((out2 - ref).norm() / ref.norm()).item() < 3e-2


def test_graphed_train_step_matches_eager():
    """GraphedTrainStep (zero_grad + forward + loss + backward + optimizer step captured once, then replayed) follows
    the eager step loop: same losses on a dropout-free net (the reference's dropout_op=None hook) over 7 steps, of which
    3 run eagerly as warm-up, one captures and 3 replay.  Covers capture of the batched weight pack / gradient unpack,
    the statistic and accumulator arenas and the fused optimizer inside the graph."""
    def build():
        torch.manual_seed(11)
        nd = {'dropout_op': None}
        return unet3d_b200.Unet(1, 3, unet3d_b200.generate_paired_features(2, 8), pool_block=unet3d_b200.ResBlock,
                                pool_kwargs={'stride': 2, **nd}, encode_block=unet3d_b200.ResBlockStack, encode_kwargs=nd,
                                encode_kwargs_fn=lambda level: {'num_stacks': max(level, 1)},
                                decode_block=unet3d_b200.ResBlock, decode_kwargs=nd).to(DEV).train()
    g = torch.Generator().manual_seed(2)
    xs = [torch.randn(2, 1, 16, 16, 16, generator=g).to(DEV) for _ in range(7)]
    ys = [torch.randint(0, 3, (2, 16, 16, 16), generator=g).to(DEV) for _ in range(7)]
    loss_fn = unet3d_b200.DiceLoss()
    m1 = build()
    o1 = torch.optim.Adam(m1.parameters(), lr=1e-3)
    eager = []
    for x, y in zip(xs, ys):
        o1.zero_grad(set_to_none=True)
        l = loss_fn(m1(x), y)
        l.backward()
        o1.step()
        eager.append(l.item())
    m2 = build()
    o2 = torch.optim.Adam(m2.parameters(), lr=1e-3)
    step = unet3d_b200.GraphedTrainStep(m2, loss_fn, o2, warmup=3)
    graphed = []
    for x, y in zip(xs, ys):
        l, _ = step(x, y)
        graphed.append(l.item())
    torch.cuda.synchronize()
    ops.check_device_errors()
    assert step.graph is not None
    print("eager  ", [round(v, 5) for v in eager])
    print("graphed", [round(v, 5) for v in graphed])
    assert all(abs(a - b) < 2e-3 for a, b in zip(eager, graphed))
    assert eager[-1] < eager[0]
    for (n1, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        # Adam normalises the update: an element whose tiny gradient changes sign with the atomic summation order moves
        # by +-lr per step in either run, so the parameters agree to a few lr x steps, not to rounding
        assert ((p1 - p2).norm() / p1.norm().clamp_min(1e-12)).item() < 0.1, n1


@pytest.mark.parametrize("fused", [False, True])
def test_cached_window_graph_follows_the_weights(fused):
    """predict_per_patch keeps the captured window forward on the engine and, when nothing changed, replays it from the
    first window of the next volume.  After a training step (fused Adam does not bump tensor versions), an in-place
    weight edit or load_state_dict the replayed result must still equal an eager, graph-free run on the new weights."""
    torch.manual_seed(0)
    model = unet3d_b200.ResUnet3D(num_pool=1, num_features=8, out_channels=3).to(DEV)
    vol = torch.randn(48, 40, 32, 1, generator=torch.Generator().manual_seed(2)).numpy()
    kw = dict(num_classes=3, patch_size=(16, 16, 16), step_per_patch=2, verbose=False, one_hot=True, window_batch=2)

    def both():
        a = unet3d_b200.predict_per_patch(vol, model, **kw)                       # cached graph / fast path
        b = unet3d_b200.predict_per_patch(vol, model, cuda_graph=False, **kw)     # eager reference
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.allclose(np.nan_to_num(a), np.nan_to_num(b), atol=1e-6), np.nanmax(np.abs(a - b))
        return a

    p0 = both()
    p1 = both()                                   # nothing changed: replay from the first window
    assert np.allclose(np.nan_to_num(p0), np.nan_to_num(p1), atol=1e-6)
    opt = torch.optim.Adam(model.parameters(), lr=5e-2, fused=fused)
    x = torch.randn(2, 1, 16, 16, 16, device=DEV)
    y = torch.randint(0, 3, (2, 16, 16, 16), device=DEV)
    model.train()
    unet3d_b200.DiceLoss()(model(x), y).backward()
    opt.step()
    model.eval()
    p2 = both()                                   # after an optimizer step
    assert np.nanmax(np.abs(p2 - p1)) > 1e-3
    with torch.no_grad():
        model.net.fc.bias[0] += 2.0               # in-place edit outside any forward (one class: softmax sees it)
    p3 = both()
    assert np.nanmax(np.abs(p3 - p2)) > 1e-3
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd["net.fc.bias"][0] -= 2.0
    model.load_state_dict(sd)
    p4 = both()
    assert np.allclose(np.nan_to_num(p4), np.nan_to_num(p2), atol=1e-5)


def test_device_prefetcher_yields_every_batch_in_order():
    host = [{"image": torch.full((2, 1, 8, 8, 8), float(i)).pin_memory(), "label": torch.full((2, 8, 8, 8), i).pin_memory(),
             "case_id": f"c{i}"} for i in range(5)]
    seen = []
    for b in unet3d_b200.DevicePrefetcher(host, DEV):
        assert b["image"].is_cuda and b["label"].is_cuda and b["label"].dtype == torch.int64
        seen.append((b["case_id"], float(b["image"].mean().item()), int(b["label"].max().item())))
    assert seen == [(f"c{i}", float(i), i) for i in range(5)]
    assert len(unet3d_b200.DevicePrefetcher(host, DEV)) == 5
