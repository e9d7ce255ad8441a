"""Multi-GPU checks, run under torchrun on one box: sliding-window inference sharded over the ranks must
return the labels AND probabilities of a single process bit for bit (fixed-point blend sums), the Trainer must keep the
ranks in lock step, and the two data-parallel loss modes:

  (i)  local Dice (default): every rank's loss on its own shard, gradients averaged == mean of the shard gradients;
  (ii) global_batch=True   : per-class sums all-reduced inside the loss, gradients SUMMED == the gradient of ONE process
                             on the concatenated batch (exact large-batch equivalence).

Both references are computed by every rank alone on the same weights.  Launched by tests/test_multi_gpu.py:
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/tools/multi_gpu_worker.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist
import unet3d_b200
from unet3d_b200 import parallel

os.environ.setdefault("NCCL_MAX_CTAS", "16")      # as bench.py: also arms the SM-limit window of the overlapped all-reduce
dist.init_process_group("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
torch.manual_seed(0)
model = unet3d_b200.ResUnet3D(num_pool=2, num_features=8, out_channels=3).to(dev).eval()      # eval: no dropout
g = torch.Generator().manual_seed(5)
x_all = torch.randn(2 * world, 1, 16, 16, 16, generator=g).to(dev)
y_all = torch.randint(0, 3, (2 * world, 16, 16, 16), generator=g).to(dev)
xs, ys = x_all[2 * rank:2 * rank + 2], y_all[2 * rank:2 * rank + 2]


def grads_of(loss_fn, x, y):
    model.zero_grad(set_to_none=True)
    loss_fn(model(x), y).backward()
    return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


ok = True
for name, glob in (("local-dice / averaged", False), ("global-dice / summed", True)):
    loss_fn = unet3d_b200.HybirdLoss(weight_v=[1, 148, 191], alpha=0.9, beta=0.1, global_batch=glob)
    model.zero_grad(set_to_none=True)
    loss_fn(model(xs), ys).backward()
    parallel.all_reduce_gradients(model, average=not glob)
    got = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    ref_fn = unet3d_b200.HybirdLoss(weight_v=[1, 148, 191], alpha=0.9, beta=0.1)
    if glob:
        ref = grads_of(ref_fn, x_all, y_all)
    else:
        parts = [grads_of(ref_fn, x_all[2 * r:2 * r + 2], y_all[2 * r:2 * r + 2]) for r in range(world)]
        ref = {n: sum(p[n] for p in parts) / world for n in parts[0]}
    errs = sorted(((rel(got[n], ref[n]), n) for n in ref if not (n.endswith("bias") and ("conv1" in n or "conv2" in n))),
                  reverse=True)
    worst = errs[0][0]
    if rank == 0 and worst > 1e-4:        # diagnostics: a uniform scale error, or noise?
        for e, n in errs[:4]:
            a, b = got[n].double().flatten(), ref[n].double().flatten()
            print(f"    {n}: rel {e:.2e}, |got|/|ref| {float(a.norm() / b.norm()):.6f}, cosine {float(a @ b / (a.norm() * b.norm())):.8f}, "
                  f"elements differing by > 1e-3 relative: {int(((a - b).abs() > 1e-3 * b.abs().max()).sum())} of {a.numel()}")
    if rank == 0:
        print(f"{name}: worst per-tensor rel-L2 vs the single-process reference {worst:.2e} ({errs[0][1]}; next {errs[1][0]:.1e} {errs[1][1]})")
    ok = ok and worst < 2e-2
# BatchNorm variant: SyncBN (statistics of the global batch) + global-batch Dice + summed gradients must equal ONE process
# on the concatenated batch, running statistics included; without SyncBN the two differ (per-rank statistics)
torch.manual_seed(1)
bn = unet3d_b200.ResAttrBNUnet3D(num_pool=1, num_features=8, out_channels=3).to(dev).train()
for mod in bn.modules():
    if isinstance(mod, torch.nn.Dropout3d):
        mod.p = 0.0                                            # no dropout: the runs must be comparable
sd0 = {k: v.clone() for k, v in bn.state_dict().items()}
loss_g = unet3d_b200.DiceLoss(global_batch=True)
parallel.enable_sync_batchnorm(bn)
bn.zero_grad(set_to_none=True)
loss_g(bn(xs), ys).backward()
parallel.all_reduce_gradients(bn, average=False)
got = {n: p.grad.detach().clone() for n, p in bn.named_parameters() if p.grad is not None}
got_buf = {n: b.detach().clone() for n, b in bn.named_buffers()}
parallel.enable_sync_batchnorm(bn, False)
bn.load_state_dict(sd0)
bn.zero_grad(set_to_none=True)
unet3d_b200.DiceLoss()(bn(x_all), y_all).backward()
ref = {n: p.grad.detach().clone() for n, p in bn.named_parameters() if p.grad is not None}
# conv biases in front of a training-mode BatchNorm have a mathematically zero gradient (the batch mean absorbs them): what
# is computed is the rounding noise of a sum of 16-bit values, ~1e-3 of the real gradients, with a relative error of ~1
# between any two runs (tools/det_probe_bn.py shows it for ONE process computing the same batch twice).  They are the
# conv1 / conv2 biases of the residual blocks (the transposed conv's bias is NOT cancelled: the zero pad planes inserted
# between it and the norm do not carry it); checked for smallness only.
import re
big = max(r.norm() for r in ref.values())
zero_grad = [n for n in ref if re.search(r"(conv1|conv2)\.bias$", n)]
assert all(got[n].norm() < 1e-2 * big and ref[n].norm() < 1e-2 * big for n in zero_grad), "BatchNorm-cancelled bias gradients are not small"
worst = max(rel(got[n], ref[n]) for n in ref if n not in zero_grad)
worst_buf = max(rel(got_buf[n].double(), b.detach().double()) for n, b in bn.named_buffers() if b.dtype.is_floating_point)
if rank == 0:
    print(f"SyncBN + global-dice / summed: worst per-tensor gradient rel-L2 vs one process on the whole batch {worst:.2e}, "
          f"running statistics {worst_buf:.2e}")
if not (worst < 2e-2 and worst_buf < 1e-5):
    print(f"[rank {rank}] SyncBN check failed: gradients {worst:.3e}, running statistics {worst_buf:.3e}", flush=True)
ok = ok and worst < 2e-2 and worst_buf < 1e-5
import numpy as np
# ---- sliding-window inference, windows dealt to the ranks: the blend sums are 2^54 fixed point (exact integer addition),
# so the sharded result must equal the single-process result BIT FOR BIT -- labels and probabilities, uniform and Gaussian
vol = np.random.RandomState(3).standard_normal((96, 72, 40, 1)).astype(np.float32)
for window in (None, "gaussian"):
    sharded = unet3d_b200.predict_per_patch(vol, model, 3, (32, 32, 32), 2, verbose=False, window=window)
    single = unet3d_b200.predict_per_patch(vol, model, 3, (32, 32, 32), 2, verbose=False, distributed=False, window=window)
    probs_s = unet3d_b200.predict_per_patch(vol, model, 3, (32, 32, 32), 2, verbose=False, one_hot=True, window=window)
    probs_1 = unet3d_b200.predict_per_patch(vol, model, 3, (32, 32, 32), 2, verbose=False, one_hot=True, distributed=False,
                                            window=window)
    same_l = bool(np.array_equal(sharded, single))
    same_p = bool(np.array_equal(probs_s.view(np.uint32), probs_1.view(np.uint32)))
    if rank == 0:
        print(f"sliding-window inference over {world} ranks vs one process ({window or 'uniform'}): labels bit-identical {same_l}, "
              f"probabilities bit-identical {same_p} (NaN border voxels: {int(np.isnan(probs_1).any(-1).sum())})")
    if not (same_l and same_p):
        print(f"[rank {rank}] inference check failed ({window}): labels {same_l}, probabilities {same_p}", flush=True)
    ok = ok and same_l and same_p

# ---- Trainer under data parallelism: 7 cases over 2 ranks, different seeds per rank -> one split, equal step
# counts, identical parameters and learning rates afterwards
class _Cases(torch.utils.data.Dataset):
    def __init__(self):
        g = torch.Generator().manual_seed(11)
        self.x = torch.randn(7, 1, 16, 16, 16, generator=g)
        self.y = (self.x[:, 0] > 0.3).long() + (self.x[:, 0] > 1.0).long()
    def __len__(self):
        return 7
    def __getitem__(self, i):
        return {"image": self.x[i], "label": self.y[i]}

torch.manual_seed(100 + rank)
np.random.seed(100 + rank)
tm = unet3d_b200.ResUnet3D(num_pool=2, num_features=8, out_channels=3).to(dev)
opt = torch.optim.Adam(tm.parameters(), lr=1e-3)
tr = unet3d_b200.Trainer(tm, opt, unet3d_b200.DiceLoss(), _Cases(), batch_size=2, dataloader_kwargs={"num_workers": 0},
                         valid_split=0.3, metrics={"dice": unet3d_b200.Dice()})
tr.fit(num_epochs=2)
flat = torch.cat([p.detach().reshape(-1) for p in tm.parameters()])
both = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(both, flat)
splits = [None] * world
dist.all_gather_object(splits, (tr.train_indices, tr.valid_indices, tr.best_result["loss"]))
same_params = all(torch.equal(both[0], b) for b in both[1:])
same_split = all(s == splits[0] for s in splits[1:])
if rank == 0:
    print(f"Trainer on {world} ranks, 7 cases: identical parameters after 2 epochs {same_params}, one split / one epoch mean {same_split}")
ok = ok and same_params and same_split
# ---- GraphedTrainStep under data parallelism: ONE CUDA graph per step with the NCCL all-reduce captured inside, 1 / world
# folded into the gradient unpack and the first gradient chunks all-reduced under the backward pass -- against the same
# steps run eagerly with parallel.all_reduce_gradients (same initial weights, same per-rank batches, eval mode: no dropout)
def _dp_run(graphed, lr):
    """lr = 0: SGD that changes nothing -> returns the all-reduced GRADIENTS of the last step (a scale error would be
    invisible behind Adam's normalisation); lr > 0: Adam, returns the parameters after 7 steps."""
    torch.manual_seed(7)
    m = unet3d_b200.ResUnet3D(num_pool=2, num_features=16, out_channels=3).to(dev).eval()
    o = torch.optim.Adam(m.parameters(), lr=lr, capturable=True) if lr > 0 else torch.optim.SGD(m.parameters(), lr=0.0)
    lf = unet3d_b200.DiceLoss()
    gg = torch.Generator().manual_seed(50 + rank)
    xb = torch.randn(2, 1, 32, 32, 32, generator=gg).to(dev)
    yb = torch.randint(0, 3, (2, 32, 32, 32), generator=gg).to(dev)
    stepper = unet3d_b200.GraphedTrainStep(m, lf, o, warmup=2) if graphed else None
    losses = []
    for _ in range(7 if lr > 0 else 5):
        if stepper is not None:
            loss = stepper(xb, yb)[0]
        else:
            o.zero_grad(set_to_none=True)
            loss = lf(m(xb), yb)
            loss.backward()
            parallel.all_reduce_gradients(m)
            o.step()
        losses.append(float(loss.item()))
    torch.cuda.synchronize()
    mode = stepper.mode if stepper is not None else "eager"
    if lr > 0:
        vec = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).clone()
        _dp_run.last_params = {n: p.detach().clone() for n, p in m.named_parameters()}
    else:
        vec = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    return vec, losses, mode

class _EvalStep(unet3d_b200.GraphedTrainStep):
    def _fwd_bwd(self, image, label):          # keep the model in eval mode (no dropout) so that both runs are comparable
        self.optimizer.zero_grad(set_to_none=True)
        logits = self.model(image)
        loss = self.loss_fn(logits, label)
        loss.backward()
        return loss, logits

_orig = unet3d_b200.GraphedTrainStep
unet3d_b200.GraphedTrainStep = _EvalStep
g_eager, _, _ = _dp_run(False, 0.0)
g_graph, _, mode = _dp_run(True, 0.0)
p_eager, l_eager, _ = _dp_run(False, 1e-3)
pe_named = _dp_run.last_params
p_eager2, _, _ = _dp_run(False, 1e-3)
p_graph, l_graph, _ = _dp_run(True, 1e-3)
if rank == 0:
    worst = sorted(((rel(_dp_run.last_params[n], pe_named[n]), n) for n in pe_named), reverse=True)[:5]
    print("Adam drift by tensor (graph vs eager):", "  ".join(f"{e:.1e} {n}" for e, n in worst), flush=True)
unet3d_b200.GraphedTrainStep = _orig
gerrs = sorted(((rel(g_graph[n], g_eager[n]), n) for n in g_eager if not (n.endswith("bias") and ("conv1" in n or "conv2" in n))),
               reverse=True)
drift = rel(p_graph, p_eager)
both = [torch.zeros_like(p_graph) for _ in range(world)]
dist.all_gather(both, p_graph)
same_ranks = all(torch.equal(both[0], b) for b in both[1:])
from unet3d_b200 import engine as _eng
if rank == 0:
    print(f"SM-limit window {_eng.NCCL_WINDOW} with NCCL_MAX_CTAS={os.environ.get('NCCL_MAX_CTAS')}")
    print(f"GraphedTrainStep ({mode}) vs eager data-parallel steps: worst per-tensor GRADIENT rel-L2 {gerrs[0][0]:.2e} ({gerrs[0][1]}; "
          f"|graph| / |eager| = {float(g_graph[gerrs[0][1]].norm() / g_eager[gerrs[0][1]].norm()):.6f}); 7 Adam steps: parameter "
          f"rel-L2 {drift:.2e} (two eager runs: {rel(p_eager2, p_eager):.2e}), last loss {l_graph[-1]:.6f} vs {l_eager[-1]:.6f}, identical parameters on all ranks {same_ranks}")
# The loss sums are reproducible, so both runs backpropagate the SAME dlogits through the 16-bit chain: the gradients
# differ only by the fp32 split-K / all-reduce summation order (~3e-7) -- that is the strict check.  Seven Adam steps
# amplify that noise (elements with |g| ~ eps move by +-lr): two identical EAGER runs drift apart just as far, so the
# parameters are only required to stay within 3x that noise floor, and the per-step losses within 1e-3.
noise = rel(p_eager2, p_eager)
ok = (ok and mode == "one-graph-dp" and gerrs[0][0] < 1e-5 and same_ranks and drift < max(3 * noise, 1e-3)
      and max(abs(a - b) for a, b in zip(l_graph, l_eager)) < 1e-3)
if not ok:
    print(f"[rank {rank}] MISMATCH", flush=True)
dist.barrier()
if rank == 0:
    print("OK" if ok else "MISMATCH")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
