"""Host-side logic on CPU: tile grids / pad-crop index math (bit-exact vs the golden vectors made from the
live reference), the C ABI library exports, state_dict layout, and the world_size-2 gloo paths."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import unet3d_b200
from oracle import unet3d_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tile_centres_match_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "tile_centres.npz"))
    for key in z.files:
        ext, p, spp = (int(v) for v in key.split("_"))
        assert np.array_equal(unet3d_b200.tile_centres(ext, p, spp), z[key].astype(np.int64)), key
    assert unet3d_b200.tile_centres(512, 128, 2).tolist() == [64, 127, 190, 253, 316, 379, 442]
    assert unet3d_b200.tile_centres(512, 128, 2, "full_cover").tolist() == [64, 128, 192, 256, 320, 384, 448]
    assert len(unet3d_b200.tile_origins((512, 512, 256), (128, 128, 128), 2)) == 147


def test_tile_grid_property():
    rng = np.random.RandomState(0)
    for _ in range(200):
        p = int(rng.choice([8, 16, 24, 32, 96, 128]))
        ext = int(rng.randint(p, 4 * p + 7))
        spp = int(rng.choice([1, 2, 4]))
        a = unet3d_b200.tile_centres(ext, p, spp)
        assert np.array_equal(a, O.tile_centres(ext, p, spp))
        assert a[0] - p // 2 == 0 and a[-1] + p // 2 <= ext
        f = unet3d_b200.tile_centres(ext, p, spp, "full_cover")
        cover = np.zeros(ext, bool)
        for c in f:
            cover[c - p // 2:c + p // 2] = True
        assert cover[:ext - (p % 2)].all()


def test_pad_crop_bit_exact():
    rng = np.random.RandomState(1)
    for shape, size in [((5, 6, 4), (9, 8, 8)), ((9, 4, 8), (4, 6, 8)), ((6, 6, 6), (6, 6, 6)), ((3, 10, 7), (8, 8, 8))]:
        a = rng.randn(*shape).astype(np.float32)
        assert np.array_equal(unet3d_b200.center_pad_crop(a, size), O.crop_pad(a, size))
        assert np.array_equal(unet3d_b200.pad_to_patch(a, size), O.pad_to(a, size))
    v = rng.randn(5, 6, 4, 2).astype(np.float32)
    assert np.array_equal(unet3d_b200.pad_to_patch(v, (8, 8, 8)), O.pad_to(v, (8, 8, 8)))


def test_gaussian_window_matches_oracle():
    assert np.array_equal(unet3d_b200.gaussian_window((8, 12, 6)), O.gaussian_window((8, 12, 6)))


def test_library_exports_every_declared_symbol():
    from unet3d_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "unet3d_b200.h")).read()
    declared = set(re.findall(r"\b(unet3d_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.exported_symbols())
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.lib().unet3d_version().startswith(b"unet3d_b200")


def test_state_dict_layout_and_unused_params():
    m = unet3d_b200.ResUnet3D(out_channels=3)
    sd = m.state_dict()
    assert len(sd) == 126 and sum(v.numel() for v in sd.values()) == 85064343
    assert sd["net.up_blocks.3.conv_trans.up.0.weight"].shape == (480, 240, 3, 3, 3)
    assert sd["net.decode_blocks.0.conv1.weight"].shape == (30, 60, 3, 3, 3)
    assert unet3d_b200.UNet3D is unet3d_b200.ResUnet3D
    assert (m.num_pool, m.num_features, m.in_channels, m.out_channels) == (4, 30, 1, 3)
    with pytest.raises(AssertionError):
        unet3d_b200.Unet(1, 1, [[8, 8], [16, 16]])          # even number of pairs (network.py:491)
    with pytest.raises(AssertionError):
        unet3d_b200.Unet(1, 1, [[8, 8]])                    # no pooling level (network.py:493)
    with pytest.raises(NotImplementedError):
        unet3d_b200.ResBlock(8, 8, norm_op=torch.nn.GroupNorm)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 16, 16, 16))         # CPU tensors are refused: no fallback path


@pytest.mark.parametrize("fixture,build", [
    ("small_resunet.npz", lambda: unet3d_b200.ResUnet3D(num_pool=2, num_features=8, out_channels=3)),
    ("attr_resunet.npz", lambda: unet3d_b200.ResAttrUnet3D(num_pool=2, num_features=8, out_channels=3)),
    ("plain_unet_train.npz", lambda: unet3d_b200.Unet(1, 3, unet3d_b200.generate_paired_features2(2, 4))),
])
def test_variants_load_reference_state_dicts(golden_dir, fixture, build):
    """Keys, shapes and -- where recorded -- the parameter ORDER (what optimizer state_dicts index by) equal the
    reference's (network.py:536-547 registers pool, up, encode, decode, conv, fc)."""
    z = np.load(os.path.join(golden_dir, fixture))
    ref = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    m = build()
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys()) or set(sd.keys()) == set(ref.keys())
    assert all(sd[k].shape == ref[k].shape for k in ref)
    m.load_state_dict(ref, strict=True)
    if "param_order" in z.files:
        assert [k for k, _ in m.named_parameters()] == z["param_order"].tolist()


def _bn_net():
    bn = {'norm_op': torch.nn.BatchNorm3d}
    nd = {'norm_op': torch.nn.BatchNorm3d, 'dropout_op': None}
    return unet3d_b200.Unet(1, 3, unet3d_b200.generate_paired_features(2, 4), pool_block=unet3d_b200.ResBlock,
                            pool_kwargs={'stride': 2, **nd}, up_kwargs={'attention': True, **bn},
                            encode_block=unet3d_b200.ResBlockStack, encode_kwargs=nd,
                            encode_kwargs_fn=lambda level: {'num_stacks': max(level, 1)},
                            decode_block=unet3d_b200.ResBlock, decode_kwargs=nd)


def test_batchnorm_variant_state_dict(golden_dir):
    """BatchNorm3d variant: the reference's kwargs hooks build the same module tree (keys incl. running buffers, shapes,
    parameter order); ResAttrBNUnet3D is that wiring with the default dropout (network.py:38-69)."""
    z = np.load(os.path.join(golden_dir, "bn_attr_resunet.npz"))
    ref = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    m = _bn_net()
    assert list(m.state_dict().keys()) == list(ref.keys())
    m.load_state_dict(ref, strict=True)
    assert [k for k, _ in m.named_parameters()] == z["param_order"].tolist()
    full = unet3d_b200.ResAttrBNUnet3D(num_pool=2, num_features=4, out_channels=3)
    assert ["net." + k for k in ref] == list(full.state_dict().keys())
    assert full.net.pool_blocks[0].dropout_p == 0.5 and isinstance(full.net.pool_blocks[0].norm, torch.nn.BatchNorm3d)


def test_default_net_parameter_order_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "default_resunet_32.npz"))
    m = unet3d_b200.ResUnet3D(out_channels=3)
    assert [k for k, _ in m.named_parameters()] == z["names"].tolist()
    assert list(m.state_dict().keys()) == z["state_keys"].tolist()
    assert unet3d_b200.generate_paired_features2(2, 4) == O.paired_features2(2, 4)
    a = unet3d_b200.ResAttrUnet3D2(out_channels=3)
    assert a.net.num_pool == 5 and a.net.up_blocks[0].att_gate.conv.weight.shape == (30, 30, 1, 1, 1)
    assert a.net.encode_blocks[5].res_blocks[0].conv1.weight.shape == (320, 320, 3, 3, 3)


def _gloo_worker(rank, world, port):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unet3d_b200 import parallel
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    unused = torch.nn.Parameter(torch.zeros(5))
    lin.register_parameter("unused", unused)
    grads = []
    for r in range(world):                     # what every rank would compute locally
        lin.zero_grad()
        lin(torch.full((2, 4), float(r + 1))).sum().backward()
        grads.append([p.grad.clone() for p in lin.parameters() if p.grad is not None])
    lin.zero_grad()
    unused.grad = None
    lin(torch.full((2, 4), float(rank + 1))).sum().backward()
    parallel.all_reduce_gradients(lin)
    got = [p.grad for p in lin.parameters() if p.grad is not None]
    for i, g in enumerate(got):
        assert torch.allclose(g, sum(gr[i] for gr in grads) / world), i
    assert unused.grad is None                 # parameters without a gradient are left alone
    buf = torch.zeros(7)
    for t in parallel.shard(list(range(7)), rank, world):
        buf[t] += t + 1
    parallel.all_reduce_sum([buf])
    assert buf.tolist() == [1, 2, 3, 4, 5, 6, 7]     # every window handled exactly once across ranks
    buf2 = torch.zeros(7)
    mine = parallel.shard_contiguous(list(range(7)), rank, world)
    assert mine == list(range(mine[0], mine[-1] + 1))              # a contiguous share
    for t in mine:
        buf2[t] += t + 1
    parallel.all_reduce_sum([buf2])
    assert buf2.tolist() == [1, 2, 3, 4, 5, 6, 7]
    assert parallel.rank_world() == (rank, world)
    dist.destroy_process_group()


def test_gloo_world2_gradient_allreduce_and_tile_sharding():
    import torch.multiprocessing as mp
    port = 29000 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(2, port), nprocs=2, join=True)


class _OddCases(torch.utils.data.Dataset):
    """7 cases: not divisible by 2 ranks, and after a 0.3 validation split neither part is."""

    def __init__(self):
        g = torch.Generator().manual_seed(11)
        self.x = torch.randn(7, 1, 4, 4, 4, generator=g)
        self.y = (self.x[:, 0] > 0).long()

    def __len__(self):
        return 7

    def __getitem__(self, i):
        return {"image": self.x[i], "label": self.y[i]}


def _gloo_trainer_worker(rank, world, port, tmp):
    """Trainer under data parallelism (trainer.py:415-604 has no multi-GPU code): every rank must see ONE
    train / valid split, run the same number of steps although 7 cases do not divide by 2, start from rank 0's
    parameters, and hand the scheduler / best-checkpoint logic the same epoch means."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import unet3d_b200
    torch.manual_seed(100 + rank)                 # different initial weights and shuffles per rank on purpose
    np.random.seed(100 + rank)
    model = torch.nn.Conv3d(1, 2, 3, padding=1)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=0.5, patience=0, threshold=10.0)
    tr = unet3d_b200.Trainer(model, opt, torch.nn.CrossEntropyLoss(), _OddCases(), batch_size=2,
                             dataloader_kwargs={"num_workers": 0}, valid_split=0.3, scheduler=sched)
    splits = [None] * world
    dist.all_gather_object(splits, (tr.train_indices, tr.valid_indices))
    assert splits[0] == splits[1]                                  # one split for the whole job
    assert sorted(splits[0][0] + splits[0][1]) == list(range(7))
    tr.fit(num_epochs=3, save_dir=os.path.join(tmp, "ck"))
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    both = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    assert torch.equal(both[0], both[1])                           # same start, same averaged gradients
    state = [None] * world
    dist.all_gather_object(state, (tr.get_lr(), tr.best_result["loss"]))
    assert state[0] == state[1]                                    # reduced epoch means: schedulers in lock step
    assert tr.get_lr() < 0.1                                       # the plateau scheduler did act
    if rank == 0:
        assert os.path.exists(os.path.join(tmp, "ck-last.pt"))
    dist.destroy_process_group()


def test_gloo_world2_trainer_odd_case_count(tmp_path):
    import torch.multiprocessing as mp
    port = 31000 + os.getpid() % 2000
    mp.spawn(_gloo_trainer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)


def test_checkpoint_loads_without_unrestricted_pickle(tmp_path):
    """Reference checkpoints hold numpy scalars in best_result (trainer.py:606-619): they must load through
    weights_only=True with the allow-list, i.e. also with allow_pickle=False."""
    model = torch.nn.Conv3d(1, 2, 1)
    opt = torch.optim.Adam(model.parameters())
    tr = unet3d_b200.Trainer(model, opt, torch.nn.CrossEntropyLoss(), _OddCases(), batch_size=2,
                             dataloader_kwargs={"num_workers": 0}, valid_split=0.0)
    tr.best_result = {"loss": np.float64(0.25), "dice": np.float32(0.5)}
    tr.current_epoch = 4
    path = str(tmp_path / "ref-like.pt")
    tr.save_checkpoint(path)
    tr2 = unet3d_b200.Trainer(torch.nn.Conv3d(1, 2, 1), torch.optim.Adam(model.parameters()), torch.nn.CrossEntropyLoss(),
                              _OddCases(), batch_size=2, dataloader_kwargs={"num_workers": 0}, valid_split=0.0)
    tr2.load_checkpoint(path, allow_pickle=False)
    assert tr2.current_epoch == 5 and float(tr2.best_result["loss"]) == 0.25
    assert tr2.train_indices == tr.train_indices


def test_sm_limit_window_bookkeeping(monkeypatch):
    """ops.set_sm_limit / _count: the limit set for N launches is dropped by the launch after the N-th; the weight
    gradient's split-K count is re-derived for the limited SM count (engine.NCCL_WINDOW)."""
    from unet3d_b200 import ops, _lib

    calls = []

    class _Fake:
        def unet3d_set_sm_limit(self, limit):
            calls.append(int(limit))
            return 0

    monkeypatch.setattr(_lib, "lib", lambda: _Fake())
    monkeypatch.setattr(ops, "SM_LIMIT", 0)
    monkeypatch.setattr(ops, "_SM_LIMIT_UNTIL", 0)
    base = ops.LAUNCHES
    ops.set_sm_limit(132, 3)
    assert calls == [132] and ops.SM_LIMIT == 132
    for _ in range(3):
        ops._count()
        assert ops.SM_LIMIT == 132
    ops._count()                                   # the 4th launch runs on all SMs again
    assert ops.SM_LIMIT == 0 and calls == [132, 0]
    ops.set_sm_limit(0)                            # idempotent: no library call
    ops.set_sm_limit(132, 0)                       # zero launches = off
    assert calls == [132, 0] and ops.SM_LIMIT == 0
    monkeypatch.setattr(ops, "LAUNCHES", base)
    # jobs x split CTAs, one per SM
    assert ops.limited_split(1, 148, 0) == 148
    assert ops.limited_split(1, 148, 132) == 132
    assert ops.limited_split(4, 37, 132) == 33
    assert ops.limited_split(36, 4, 132) == 3
    assert ops.limited_split(132, 1, 132) == 1
    assert ops.limited_split(160, 1, 132) == 1      # more jobs than SMs: several waves either way
    assert ops.limited_split(9, 14, 132) == 14      # 126 CTAs already fit
