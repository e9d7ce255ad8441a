"""Drop-in replacements for the hot-path parts of the reference's ``trainer.py``:

  predict_per_patch(input, model, num_classes=3, patch_size=(96,96,96), step_per_patch=4,
                    verbose=True, one_hot=False)                                   trainer.py:17-98
  predict_case(case, model, target_spacing, normalize_stats, num_classes=3, patch_size=(96,96,96),
               step_per_patch=4, verbose=True, one_hot=False)                      trainer.py:101-133
  cascade_predict_case(case, coarse_model, ..., detail_model, ..., num_classes=3, step_per_patch=4,
                       region_threshold=10000, crop_padding=20, verbose=True)      trainer.py:164-245
  Trainer(model, optimizer, loss, dataset, ...).fit / batch_loop / save_checkpoint / load_checkpoint
                                                                                   trainer.py:415-634

Same arguments, same return values, same tile grid (including the float-step truncation of
trainer.py:34-40) and the same checkpoint dictionary.  New, optional capabilities the
reference does not have: Gaussian-weighted blending (``window="gaussian"``), a ``grid_mode="full_cover"``
tile grid that reaches the volume border, tiles of one volume sharded over the ranks of a
``torch.distributed`` job, and data-parallel training with a gradient all-reduce.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from . import parallel


# ------------------------------------------------------------------------------------------------
# integer index math (bit-exact with transform.py:387-437 / trainer.py:29-65)
# ------------------------------------------------------------------------------------------------
def center_pad_crop(a: np.ndarray, size: Sequence[int], cval=0) -> np.ndarray:
    """Centre crop-or-pad of the leading len(size) axes to `size` (transform.py:393-437, centre mode).
    The low edge of the window is floor((extent - size) / 2); missing voxels are filled with cval."""
    out_shape = list(a.shape)
    src, dst = [], []
    for d, s in enumerate(size):
        lo = (a.shape[d] - s) // 2                  # may be negative: then -lo voxels of padding in front
        hi = lo + s
        s0, s1 = max(lo, 0), min(hi, a.shape[d])
        src.append(slice(s0, s1))
        dst.append(slice(s0 - lo, s1 - lo))
        out_shape[d] = s
    out = np.full(out_shape, cval, dtype=a.dtype)
    out[tuple(dst)] = a[tuple(src)]
    return out


def pad_to_patch(a: np.ndarray, patch: Sequence[int]) -> np.ndarray:
    """transform.py:387-390: grow every axis that is shorter than the patch."""
    return center_pad_crop(a, [max(a.shape[d], patch[d]) for d in range(len(patch))])


def tile_centres(extent: int, patch: int, step_per_patch: int, mode: str = "reference") -> np.ndarray:
    """Window centres along one axis.

    mode="reference": exactly what ``np.arange(start, end + 1e-8, step, dtype=np.int)`` produced in the
    reference (trainer.py:29-40): the float step is a hair below the nominal stride, numpy derives the
    integer stride from the first two elements, so the stride is one voxel short and the last
    voxels of the axis may stay uncovered (they come out as label 0 / NaN probabilities).
    mode="full_cover": nominal stride, last window clamped to the border."""
    start = patch // 2
    end = extent - patch // 2
    n_steps = math.ceil((end - start) / (patch / step_per_patch))
    if mode == "reference":
        step = (end - start) / (n_steps + 1e-8)
        if step == 0:
            step = 9999999
        count = int(math.ceil((end + 1e-8 - start) / step))
        delta = int(start + step) - start
        return start + delta * np.arange(count, dtype=np.int64)
    if mode == "full_cover":
        if n_steps == 0:
            return np.array([start], dtype=np.int64)
        return np.unique(np.round(start + (end - start) * np.arange(n_steps + 1) / n_steps).astype(np.int64))
    raise ValueError(mode)


def tile_origins(shape: Sequence[int], patch: Sequence[int], step_per_patch: int, mode: str = "reference"):
    """Window origins in the reference's visiting order (x outermost, z fastest; trainer.py:53-65)."""
    cs = [tile_centres(shape[d], patch[d], step_per_patch, mode) for d in range(3)]
    return [(int(x) - patch[0] // 2, int(y) - patch[1] // 2, int(z) - patch[2] // 2)
            for x in cs[0] for y in cs[1] for z in cs[2]]


def gaussian_window(patch: Sequence[int], sigma_scale: float = 0.125) -> np.ndarray:
    """Separable Gaussian importance map, w(i) = exp(-((i - (p-1)/2) / (sigma_scale p))^2 / 2)."""
    ax = [np.exp(-0.5 * ((np.arange(p, dtype=np.float64) - (p - 1) / 2.0) / (sigma_scale * p)) ** 2) for p in patch]
    return (ax[0][:, None, None] * ax[1][None, :, None] * ax[2][None, None, :]).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# sliding-window inference
# ------------------------------------------------------------------------------------------------
def predict_per_patch(input, model, num_classes=3, patch_size=(96, 96, 96), step_per_patch=4, verbose=True,
                      one_hot=False, window=None, grid_mode="reference", window_batch=4, cuda_graph=True,
                      distributed=True):
    """input: (X, Y, Z, C_in) float32 numpy.  Returns uint8 labels (X, Y, Z) or, with one_hot=True,
    float32 probabilities (X, Y, Z, num_classes) -- trainer.py:17-98.

    window: None = uniform blending (the reference), "gaussian" or a (px,py,pz) float array = weighted.
    window_batch: windows per forward pass (the reference runs one; InstanceNorm and the blend are per window, so
    the result does not depend on it -- it amortises the ~300 kernel launches of a forward pass and the weight
    traffic of the deep, weight-bound levels: 0.567 / 0.517 / 0.503 s per 512x512x256 volume at 2 / 4 / 8).
    cuda_graph: capture the forward pass of one window batch once and replay it for the others (this library's
    models only; a window forward is ~300 kernel launches and is host-bound when launched eagerly).
    Under an initialised torch.distributed job the window list is cut into contiguous shares (x-slabs), each rank
    uploads only its slab, and the partial sums are all-reduced; every rank returns the full result."""
    import os
    import time
    dbg = os.environ.get("U3D_PREDICT_TIMES")
    t_dbg = [time.perf_counter()]

    def mark(what):
        if dbg:
            torch.cuda.synchronize()
            t_dbg.append(time.perf_counter())
            print(f"[predict_per_patch rank {parallel.rank_world()[0]}] {what}: {t_dbg[-1] - t_dbg[-2]:.3f} s", flush=True)

    patch = tuple(int(p) for p in patch_size)
    orig_shape = tuple(input.shape[:3])
    vol = np.asarray(input, dtype=np.float32)
    if any(vol.shape[d] < patch[d] for d in range(3)):       # transform.pad: only volumes smaller than a patch are copied
        vol = pad_to_patch(vol, patch)
    pshape = tuple(vol.shape[:3])
    if vol.shape[3] == 1:
        host = np.ascontiguousarray(vol).reshape(1, 1, *pshape)                                  # (1, 1, X, Y, Z): a view
    else:
        host = np.ascontiguousarray(np.moveaxis(vol, -1, 0))[None]                               # (1, C, X, Y, Z)

    def upload(mine, device):
        """The device copy of the volume is filled x-slab by x-slab, on a side stream, just ahead of the windows that read
        it: the host-side staging of the (pageable) numpy volume overlaps the forward passes already enqueued, and a rank
        of a sharded run only ever uploads the slabs of its own windows."""
        x = torch.empty(host.shape, dtype=torch.float32, device=device)
        side = torch.cuda.Stream(device=device)
        state = {"lo": None, "hi": None}

        def ensure(xlo, xhi):
            if state["lo"] is None:
                todo = [(xlo, xhi)]
                state["lo"], state["hi"] = xlo, xhi
            else:
                todo = []
                if xlo < state["lo"]:
                    todo.append((xlo, state["lo"]))
                    state["lo"] = xlo
                if xhi > state["hi"]:
                    todo.append((state["hi"], xhi))
                    state["hi"] = xhi
            if not todo:
                return
            with torch.cuda.stream(side):
                for a, b in todo:
                    x[:, :, a:b].copy_(torch.from_numpy(host[:, :, a:b]), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
            torch.cuda.current_stream(device).wait_event(ev)
        return x, ensure

    with torch.cuda.device(next(model.parameters()).device):       # launches go to the current device's stream
        res = _predict_padded(upload, pshape, model, num_classes, patch, step_per_patch, verbose, one_hot, window,
                              grid_mode, window_batch, cuda_graph, distributed, mark)
    # _predict_padded returns with the GPU still working: the host result buffer is allocated and touched (page faults:
    # ~25 ms for a 512x512x256 label map) while the windows run, so the copy at the end only moves bytes.  (A pinned
    # buffer would copy faster but costs ~100 ms of cudaHostAlloc whenever PyTorch's pinned cache has no free block.)
    out = torch.empty(res.shape, dtype=res.dtype)
    out.zero_()
    out.copy_(res)
    out = out.numpy()
    ops.check_device_errors()
    mark("D2H")
    return out if pshape == orig_shape else center_pad_crop(out, orig_shape)


def _predict_padded(upload, shape, model, num_classes, patch, step_per_patch, verbose, one_hot, window, grid_mode,
                    window_batch, cuda_graph, distributed, mark=lambda what: None):
    """The window loop on a volume that is already padded to at least one patch (trainer.py:43-96).
    upload(windows of this rank, device) -> (1, C, X, Y, Z) float32 device tensor.  Returns a DEVICE tensor on the padded
    grid: uint8 labels (X, Y, Z) or float32 probabilities (X, Y, Z, num_classes)."""
    device = next(model.parameters()).device
    shape = tuple(int(v) for v in shape)
    origins = tile_origins(shape, patch, step_per_patch, grid_mode)
    rank, world = parallel.rank_world() if distributed else (0, 1)      # distributed=False: this process alone
    mine = parallel.shard_contiguous(origins, rank, world)          # an x-slab of the volume per rank
    x = upload(mine, device)
    ensure = None
    if isinstance(x, tuple):               # (device volume, ensure(xlo, xhi)): slabs are uploaded on demand
        x, ensure = x
    # 2^54 fixed-point sums (exact, order-independent: see ops.sw_accumulate), one x-slab per rank; channel K = weight sum
    xs_len = -(-shape[0] // world)
    acc = torch.zeros((world, num_classes + 1, xs_len, shape[1], shape[2]), dtype=torch.int64, device=device)
    if isinstance(window, str):
        if window != "gaussian":
            raise ValueError(window)
        window = gaussian_window(patch)
    wdev = None if window is None else torch.as_tensor(window, dtype=torch.float32, device=device).contiguous()

    mark("buffers (+ H2D when the volume is not uploaded slab by slab)")
    was_training = model.training
    model.eval()
    it = None
    if verbose:
        from tqdm import tqdm
        it = tqdm(total=len(mine))
    wb = max(1, int(window_batch))
    from . import network as _nw
    graphable = cuda_graph and isinstance(model, (_nw.Unet, _nw._ResNetBase)) and len(mine) > 2 * wb
    gkey = (wb, x.shape[1], patch, getattr(model, "precision", "bf16"), ops.BIAS_EPOCH)
    static_in = torch.empty((wb, x.shape[1], *patch), dtype=torch.float32, device=device)
    graph, static_out = None, None
    first = True
    eng = getattr(model.net if hasattr(model, "net") else model, "_engine", None) if graphable else None
    if eng is not None and eng.device == device and gkey in eng._infer_graphs and eng.replay_is_current():
        # a forward of this shape was captured by an earlier call and neither the weights nor the mode changed since:
        # replay from the first window on -- no eager forward (~300 Python-side launches, ~30 ms) per volume
        graph, static_in, static_out = eng._infer_graphs[gkey]
        first = False
    with torch.no_grad():
        for b0 in range(0, len(mine), wb):
            group = mine[b0:b0 + wb]
            # a short last group is padded with its first window so that only ONE batch shape is ever planned
            padded = group + [group[0]] * (wb - len(group))
            if ensure is not None:
                ensure(min(o[0] for o in group), max(o[0] for o in group) + patch[0])
            for i, (ox, oy, oz) in enumerate(padded):
                static_in[i].copy_(x[0, :, ox:ox + patch[0], oy:oy + patch[1], oz:oz + patch[2]])
            if graph is not None:
                graph.replay()
                logits = static_out
            else:
                logits = model(static_in)                    # eager: builds the plans, (re)packs the weights
                if graphable and first:
                    # the captured forward is kept on the engine and reused by later calls (other volumes): capture
                    # + instantiation of ~200 launches costs more than one volume's worth of host launch overhead
                    eng = (model.net if hasattr(model, "net") else model)._engine
                    gkey = gkey[:-1] + (ops.BIAS_EPOCH,)        # the eager forward above may have re-packed biases
                    cached = eng._infer_graphs.get(gkey)
                    if cached is None:
                        torch.cuda.synchronize()
                        ops.check_device_errors()
                        g_in = torch.empty_like(static_in)
                        g_in.copy_(static_in)
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            g_out = model(g_in)
                        for stale in [k for k in eng._infer_graphs if k[:-1] == gkey[:-1]]:
                            del eng._infer_graphs[stale]         # same shape, older bias epoch: its memory pool goes too
                        cached = eng._infer_graphs[gkey] = (g, g_in, g_out)
                    graph, static_in_new, static_out = cached
                    static_in = static_in_new
            first = False
            for i, origin in enumerate(group):
                ops.sw_accumulate(logits[i].contiguous(), wdev, acc, origin, shape)
            if verbose:
                it.update(len(group))
            if b0 < 6 * wb:
                mark(f"  group {b0 // wb} ({'replay' if graph is not None and b0 else 'eager'})")
    if it is not None:
        it.close()
    model.train(was_training)
    mark(f"{len(mine)} windows")
    if world > 1:
        # every rank ends up with the complete sums of ITS x-slab (int64 addition: any reduction order gives the same
        # bits), finalises that slab, and the uint8 label slabs (or probability slabs) are gathered
        import torch.distributed as dist
        mine = torch.empty(acc.shape[1:], dtype=torch.int64, device=device)
        dist.reduce_scatter_tensor(mine, acc)
        mark("reduce-scatter of the blend sums")
    else:
        mine = acc[0]
    slab_shape = (xs_len, shape[1], shape[2])
    if one_hot:
        part = torch.empty((*slab_shape, num_classes), dtype=torch.float32, device=device)
        ops.sw_finalize(mine, None, part)
    else:
        part = torch.empty(slab_shape, dtype=torch.uint8, device=device)
        ops.sw_finalize(mine, part, None)
    if world > 1:
        full = torch.empty((world, *part.shape), dtype=part.dtype, device=device)
        dist.all_gather_into_tensor(full, part)
        res = full.view(world * xs_len, *part.shape[1:])[:shape[0]]
        if not res.is_contiguous():
            res = res.contiguous()
    else:
        res = part
    mark("finalize")
    return res


def predict_case(case, model, *args, **kwargs):
    """trainer.py:101-133 -- see _predict_case for the arguments; runs with the model's device current."""
    with torch.cuda.device(next(model.parameters()).device):
        return _predict_case(case, model, *args, **kwargs)


def _predict_case(case, model, target_spacing, normalize_stats, num_classes=3, patch_size=(96, 96, 96),
                  step_per_patch=4, verbose=True, one_hot=False, window=None, grid_mode="reference", window_batch=4,
                  cuda_graph=True, distributed=True, keep_on_device=False):
    """trainer.py:101-133: resample + normalise the case, predict it window by window, resize the prediction back to the
    original grid.  Same arguments and result (``case['pred']``: uint8 labels, or float32 probabilities with one_hot).

    The whole chain stays in HBM: the raw image is uploaded once, ``zoom + clip + z-score`` writes straight into the
    interior of the zero-padded NCDHW model input (transform.py:387-390 pad), the label map is centre-cropped as a view
    and zoomed back by the label kernel, and only the final prediction comes back over PCIe."""
    from . import transform as T
    device = next(model.parameters()).device
    patch = tuple(int(p) for p in patch_size)
    image = case['image']                                  # numpy (X, Y, Z, C) like the reference, or already in HBM
    if not isinstance(image, torch.Tensor):
        image = np.ascontiguousarray(image, dtype=np.float32)
    orig_shape = tuple(image.shape[:-1])
    affine = case['affine']
    scale = np.array(T.get_spacing(affine)) / np.array(target_spacing)
    table = T.normalize_table(normalize_stats)
    if len(table) != image.shape[-1]:
        raise ValueError("one normalize_stats entry per image channel")
    if verbose:
        print('Resampling the case for prediction...')
    raw = image.to(device, torch.float32) if isinstance(image, torch.Tensor) else torch.from_numpy(image).to(device)
    rshape = T.zoomed_shape(orig_shape, scale)                               # the resampled grid
    pshape = tuple(max(rshape[d], patch[d]) for d in range(3))               # ... padded to at least one patch
    lo = [-((rshape[d] - pshape[d]) // 2) for d in range(3)]                 # centre pad: floor((n - size) / 2) in front
    x = torch.zeros((1, image.shape[-1], *pshape), dtype=torch.float32, device=device)
    interior = x[0, :, lo[0]:lo[0] + rshape[0], lo[1]:lo[1] + rshape[1], lo[2]:lo[2] + rshape[2]].permute(1, 2, 3, 0)
    T.rescale_device(raw, scale, multi_class=True, out=interior, norm=table)
    if verbose:
        print('Predicting the case...')
    pred = _predict_padded(lambda mine, dev: x, pshape, model, num_classes, patch, step_per_patch, verbose, one_hot,
                           window, grid_mode, window_batch, cuda_graph, distributed)
    # crop_pad back (trainer.py:98): a view.  The reference pads with ceil((p - n) / 2) voxels in front but crops
    # floor((p - n) / 2) away, so for an odd difference the result is one voxel off -- kept, it is the reference's output.
    cl = [(pshape[d] - rshape[d]) // 2 for d in range(3)]
    pred = pred[cl[0]:cl[0] + rshape[0], cl[1]:cl[1] + rshape[1], cl[2]:cl[2] + rshape[2]]
    if verbose:
        print('Resizing the case to origial shape...')
    back = np.array(orig_shape) / np.array(rshape)
    if one_hot:
        out = T.rescale_device(pred, back, multi_class=True)
    else:
        out = T.rescale_device(pred, back, is_label=True)
    case['pred'] = out if keep_on_device else out.cpu().numpy()
    case['affine'] = affine
    ops.check_device_errors()
    if verbose:
        print('All done!')
    return case


def cascade_predict_case(case, coarse_model, coarse_target_spacing, coarse_normalize_stats, coarse_patch_size,
                         detail_model, *args, **kwargs):
    """trainer.py:164-245 -- see _cascade_predict_case for the arguments; runs with the detail model's device current
    (both models must live on the same device)."""
    dev = next(detail_model.parameters()).device
    if next(coarse_model.parameters()).device != dev:
        raise RuntimeError("coarse and detail model must be on the same device")
    with torch.cuda.device(dev):
        return _cascade_predict_case(case, coarse_model, coarse_target_spacing, coarse_normalize_stats, coarse_patch_size,
                                     detail_model, *args, **kwargs)


def _cascade_predict_case(case, coarse_model, coarse_target_spacing, coarse_normalize_stats, coarse_patch_size,
                          detail_model, detail_target_spacing, detail_normalize_stats, detail_patch_size, num_classes=3,
                          step_per_patch=4, region_threshold=10000, crop_padding=20, verbose=True, **predict_kwargs):
    """trainer.py:164-245: coarse one-class prediction -> connected regions of at least ``region_threshold`` voxels, each
    grown by ``crop_padding`` mm -> detail prediction per region -> mean of the overlapping probabilities -> labels.

    Same arguments and result (``case['pred']`` uint8 on the original grid).  The image is uploaded once; the coarse
    mask, the component labelling (csrc/regions.cu), the region crops, their predictions and the merge buffers all stay
    in HBM, and only the final label map is copied back."""
    from . import transform as T
    device = next(detail_model.parameters()).device
    if verbose:
        print('Predicting the rough shape for further prediction...')
    image = case['image']
    dev_image = image if isinstance(image, torch.Tensor) else \
        torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).to(device)
    work = dict(case)
    work['image'] = dev_image
    work = predict_case(work, coarse_model, coarse_target_spacing, coarse_normalize_stats, 1, coarse_patch_size,
                        step_per_patch, verbose=verbose, keep_on_device=True, **predict_kwargs)
    work.setdefault('case_id', 'case')
    regions = T.regions_crop_case(work, region_threshold, crop_padding, 'pred')
    num_classes = detail_model.out_channels
    orig_shape = tuple(dev_image.shape[:-1])
    result = torch.zeros((*orig_shape, num_classes), dtype=torch.float64, device=device)
    result_n = torch.zeros(orig_shape, dtype=torch.int32, device=device)
    if verbose:
        print('Cropping regions (%d)...' % len(regions))
    for idx, region in enumerate(regions):
        bbox = region['bbox']
        shape = tuple(region['image'].shape[:-1])
        if verbose:
            print('Region {} {} predicting...'.format(idx, shape))
        region = predict_case(region, detail_model, detail_target_spacing, detail_normalize_stats, num_classes,
                              detail_patch_size, step_per_patch, verbose=verbose, one_hot=True, keep_on_device=True,
                              **predict_kwargs)
        src0 = [max(0 - int(bbox[i][0]), 0) for i in range(3)]
        dst0 = [max(int(bbox[i][0]), 0) for i in range(3)]
        box = [min(int(bbox[i][1]), orig_shape[i]) - dst0[i] for i in range(3)]
        ops.region_accumulate(region['pred'], result, result_n, src0, dst0, box)
    if verbose:
        print('Merging all regions...')
    case['pred'] = ops.merge_finalize(result, result_n).cpu().numpy()
    ops.check_device_errors()
    if verbose:
        print('All done!')
    return case


# ------------------------------------------------------------------------------------------------
# evaluation (trainer.py:348-356)
# ------------------------------------------------------------------------------------------------
def evaluate_case(case):
    """trainer.py:348-356: Dice (loss.dice, alpha = beta = 0.5, smooth = 1e-7) of every label 1 .. label.max() between
    ``case['pred']`` and ``case['label']`` (uint8 volumes: numpy arrays or CUDA tensors).  Returns the list of floats
    the reference returns.

    One device pass counts |pred == c and label == c|, |pred == c|, |label == c| for all labels at once (exact
    integers); the reference builds two float32 masks per label and sums them in fp32 on the CPU.  The final ratio is
    evaluated in float32 in the reference's operation order, so the result is bit-identical as long as the counts are
    exactly representable in float32 (< 2^24 voxels per label; above that the reference's own fp32 sums round)."""
    pred, label = case['pred'], case['label']
    dev = pred.device if isinstance(pred, torch.Tensor) and pred.is_cuda else (
        label.device if isinstance(label, torch.Tensor) and label.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    as_dev = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(dev).to(torch.uint8)
    counts = ops.overlap_counts(as_dev(pred), as_dev(label)).cpu().numpy()
    ops.check_device_errors()
    num_classes = int(np.nonzero(counts[2])[0].max()) if counts[2].any() else 0          # case['label'].max()
    f32 = np.float32
    out = []
    for c in range(1, num_classes + 1):
        tp = f32(counts[0, c])
        fn = f32(counts[2, c] - counts[0, c])
        fp = f32(counts[1, c] - counts[0, c])
        # loss.py:32-48: (tp + smooth) / (tp + alpha * fn + beta * fp + smooth), float32 tensors with Python-float scalars
        num = f32(tp + f32(1e-7))
        den = f32(f32(f32(tp + f32(f32(0.5) * fn)) + f32(f32(0.5) * fp)) + f32(1e-7))
        out.append(float(f32(num / den)))
    return out


# ------------------------------------------------------------------------------------------------
# training step loop
# ------------------------------------------------------------------------------------------------
class DevicePrefetcher:
    """Iterates over a loader of dict batches (trainer.py:471-473: ``batch['image'].to(device)``) and yields them on the
    device: the host->device copies of batch i+1 are enqueued on a side stream BEFORE batch i is handed out, so with
    pinned batches (the reference's DataLoader default, trainer.py:422) they run under the training step of batch i
    instead of in front of batch i+1.  Non-tensor entries pass through.

    The device tensors are two alternating sets of staging buffers per (key, shape, dtype) owned by THIS prefetcher (no
    allocator traffic in the loop; two live prefetchers never share staging memory): a batch stays valid until the
    batch after the next one is requested -- keep a ``.clone()`` if you need it longer.  Pass ``buffers=`` (a dict) to
    let several short-lived prefetchers of one owner (a Trainer's epochs) reuse the same staging sets."""

    _streams: Dict[torch.device, torch.cuda.Stream] = {}

    def __init__(self, loader, device, buffers: Optional[dict] = None):
        self.loader, self.device = loader, torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._buffers: Dict[tuple, List[torch.Tensor]] = {} if buffers is None else buffers

    def __len__(self):
        return len(self.loader)

    def _staging(self, key, t, slot):
        k = (self.device, key, tuple(t.shape), t.dtype)
        if k not in self._buffers:
            self._buffers[k] = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for _ in range(2)]
        return self._buffers[k][slot]

    def __iter__(self):
        if self.device not in self._streams:
            self._streams[self.device] = torch.cuda.Stream(device=self.device)
        side = self._streams[self.device]
        it = iter(self.loader)
        count = [0]

        def load():
            try:
                batch = next(it)
            except StopIteration:
                return None
            slot = count[0] & 1
            count[0] += 1
            # the staging set about to be overwritten was last read by the batch before the current one; all of that
            # batch's work is already enqueued on the main stream
            free = torch.cuda.Event()
            free.record(torch.cuda.current_stream(self.device))
            side.wait_event(free)
            with torch.cuda.stream(side):
                out = {k: (self._staging(k, v, slot).copy_(v, non_blocking=True) if torch.is_tensor(v) else v)
                       for k, v in batch.items()}
            ev = torch.cuda.Event()
            ev.record(side)
            return out, ev

        nxt = load()
        while nxt is not None:
            cur, ev = nxt
            torch.cuda.current_stream(self.device).wait_event(ev)
            nxt = load()          # the next upload overlaps whatever the caller enqueues for `cur`
            yield cur


class _TransformedSubset(torch.utils.data.Dataset):
    def __init__(self, dataset, indices, transform):
        self.dataset, self.indices, self.transform = dataset, list(indices), transform

    def __len__(self):
        return len(self.indices)

    def __getitem__(self, i):
        case = self.dataset[self.indices[i]]
        return self.transform(case) if self.transform else case


class Trainer:
    """Same constructor and methods as the reference's Trainer (trainer.py:415-634).

    ``use_amp=True`` in ``fit`` is accepted for compatibility: the model already computes in 16-bit
    on the tensor cores with fp32 master weights, so no apex is involved."""

    def __init__(self, model, optimizer, loss, dataset, batch_size=10,
                 dataloader_kwargs={'num_workers': 2, 'pin_memory': True}, valid_split=0.2, num_samples=None,
                 metrics=None, scheduler=None, train_transform=None, valid_transform=None, cuda_graph=False):
        self.model, self.optimizer, self.loss, self.dataset = model, optimizer, loss, dataset
        self.metrics, self.scheduler = metrics, scheduler
        self.train_transform, self.valid_transform = train_transform, valid_transform
        order = list(range(len(dataset)))
        n_valid = int(np.floor(valid_split * len(dataset)))
        np.random.shuffle(order)
        # data parallel: ONE train / valid split for the whole job (every process shuffles with its own RNG state)
        order = parallel.broadcast_object(order)
        self.train_indices, self.valid_indices = order[n_valid:], order[:n_valid]
        self.dataloader_kwargs = {'batch_size': batch_size, **dataloader_kwargs}
        self.num_samples, self.valid_split = num_samples, valid_split
        self.device = next(model.parameters()).device
        self.best_result = {'loss': float('inf')}
        self.current_epoch = 0
        self.patience_counter = 0
        self.amp_state_dict = None
        self.num_epochs, self.use_amp, self.save_dir, self.progress_bar = 1, False, None, None
        # optional (not in the reference): replay the whole training step as one CUDA graph
        self._graphed = None
        self._staging = {}                  # DevicePrefetcher staging sets, reused across epochs
        if cuda_graph:
            net = model.net if hasattr(model, "net") else model
            if parallel.rank_world()[1] > 1 and (getattr(loss, "global_batch", False) or getattr(net, "sync_bn", False)):
                raise ValueError("cuda_graph=True cannot be combined with global_batch losses or SyncBN: their "
                                 "collectives are issued from Python between kernel launches and cannot be captured")
            from .graph import GraphedTrainStep
            self._graphed = GraphedTrainStep(model, loss, optimizer)

    def get_lr(self, idx=0):
        return self.optimizer.param_groups[idx]['lr']

    def set_lr(self, lr, idx=0):
        self.optimizer.param_groups[idx]['lr'] = lr

    def summary(self, input_shape):
        n = sum(p.numel() for p in self.model.parameters())
        return f"{type(self.model).__name__}: {n} parameters, input {tuple(input_shape)}"

    # -- one pass over a loader: the hot loop (trainer.py:465-523)
    def batch_loop(self, data_loader, is_train=True):
        results = []
        bar = self.progress_bar
        if bar is not None:
            bar.reset(len(data_loader))
            bar.set_description("Epoch %d/%d (LR %.2g)" % (self.current_epoch + 1, self.num_epochs, self.get_lr()))
        batches = DevicePrefetcher(data_loader, self.device, self._staging) if self.device.type == "cuda" else data_loader
        # global-batch losses all-reduce their partial sums themselves and need SUMMED gradients (SURVEY.md 8e mode ii)
        average = not getattr(self.loss, 'global_batch', False)
        for batch in batches:
            if is_train and self._graphed is not None:
                loss, y_pred = self._graphed(batch['image'], batch['label'])
                y = self._graphed.last_label        # the label tensor the step actually used (static under replay)
                result = {'loss': loss.item()}
                if self.metrics is not None:
                    with torch.no_grad():
                        for key, fn in self.metrics.items():
                            result[key] = fn(y_pred.detach(), y).item()
                if not math.isnan(result['loss']):
                    results.append(result)
                if bar is not None:
                    bar.set_postfix(result)
                    bar.update()
                continue
            x = batch['image'].to(self.device, non_blocking=True)
            y = batch['label'].to(self.device, non_blocking=True)
            if is_train:
                self.model.train()
                y_pred = self.model(x)
            else:
                self.model.eval()
                with torch.no_grad():
                    y_pred = self.model(x)
            loss = self.loss(y_pred, y)
            if is_train:
                self.optimizer.zero_grad()
                loss.backward()
                parallel.all_reduce_gradients(self.model, average=average)
                self.optimizer.step()
            result = {'loss': loss.item()}
            if self.metrics is not None:
                with torch.no_grad():
                    for key, fn in self.metrics.items():
                        result[key] = fn(y_pred.detach(), y).item()
            if not math.isnan(result['loss']):            # NaN steps are left out of the epoch mean (trainer.py:505)
                results.append(result)
            if bar is not None:
                bar.set_postfix(result)
                bar.update()
        ops.check_device_errors()
        keys = list(results[0]) if results else ['loss'] + list(self.metrics or {})
        if parallel.rank_world()[1] > 1:
            # one epoch mean for the whole job: ReduceLROnPlateau and the best-checkpoint test must not diverge
            mean_result = parallel.all_reduce_mean_results({k: sum(r[k] for r in results) for k in keys}, len(results),
                                                           self.device)
        else:
            mean_result = {k: float(np.mean([r[k] for r in results])) for k in keys}
        if self.save_dir is not None and parallel.rank_world()[0] == 0:
            try:
                from torch.utils.tensorboard import SummaryWriter
                w = SummaryWriter(self.save_dir)
                for k, v in mean_result.items():
                    w.add_scalar('%s/%s' % (k, 'train' if is_train else 'valid'), v, self.current_epoch)
                w.close()
            except Exception:        # tensorboard is optional here
                pass
        return mean_result

    def _loader(self, indices, transform, n_samples, shuffle):
        ds = _TransformedSubset(self.dataset, indices, transform)
        rank, world = parallel.rank_world()
        if world > 1:
            # shard the cases over the ranks with EQUAL lengths (the remainder of an indivisible count is dropped): a rank
            # with one batch more would wait forever in its last gradient all-reduce
            per = len(ds) // world
            if per == 0:
                raise ValueError(f"{len(ds)} cases cannot be sharded over {world} ranks")
            ds = _TransformedSubset(ds, list(range(rank, per * world, world)), None)
        if n_samples is not None:
            sampler = torch.utils.data.RandomSampler(ds, True, max(1, n_samples // world))
            return torch.utils.data.DataLoader(ds, sampler=sampler, **self.dataloader_kwargs)
        return torch.utils.data.DataLoader(ds, shuffle=shuffle, **self.dataloader_kwargs)

    def fit(self, num_epochs=10, save_dir=None, use_amp=False, opt_level='O1'):
        from tqdm import tqdm
        self.num_epochs, self.use_amp, self.save_dir = num_epochs, use_amp, save_dir
        if use_amp and hasattr(self.model, "precision"):
            # apex O1 in the reference = fp16 activations with fp32 master weights and a dynamic loss scale
            # (trainer.py:538-542); the equivalent here is fp16 storage of activations AND gradients with the engine's
            # internal per-step power-of-two gradient scale (engine.backward_impl)
            self.model.precision = "fp16"
        self.progress_bar = tqdm(total=0, disable=parallel.rank_world()[0] != 0)
        parallel.broadcast_parameters(self.model)          # data parallel: every rank starts from rank 0's weights
        train_loader = self._loader(self.train_indices, self.train_transform, self.num_samples, True)
        valid_loader = None
        if len(self.valid_indices) > 0:
            nv = None if self.num_samples is None else round(self.num_samples * self.valid_split)
            valid_loader = self._loader(self.valid_indices, self.valid_transform, nv, False)
        for epoch in range(self.current_epoch, num_epochs):
            self.current_epoch = epoch
            result = self.batch_loop(train_loader, is_train=True)
            if valid_loader is not None:
                result = self.batch_loop(valid_loader, is_train=False)
            if self.scheduler is not None:
                if isinstance(self.scheduler, torch.optim.lr_scheduler.ReduceLROnPlateau):
                    self.scheduler.step(result['loss'])
                else:
                    self.scheduler.step()
            if result['loss'] < self.best_result['loss'] - 1e-3:
                self.best_result = result
                if save_dir is not None:
                    self.save_checkpoint(save_dir + '-best.pt')
            if save_dir is not None:
                self.save_checkpoint(save_dir + '-last.pt')
        self.progress_bar.close()

    def save_checkpoint(self, file_path):
        if parallel.rank_world()[0] != 0:
            return
        ckpt = {'model_state_dict': self.model.state_dict(), 'optimizer_state_dict': self.optimizer.state_dict(),
                'current_epoch': self.current_epoch, 'train_indices': self.train_indices,
                'valid_indices': self.valid_indices, 'best_result': self.best_result}
        if self.scheduler is not None:
            ckpt['scheduler_state_dict'] = self.scheduler.state_dict()
        torch.save(ckpt, file_path)

    def load_checkpoint(self, file_path, allow_pickle=True):
        """trainer.py:621-634.  The file is first read with ``weights_only=True`` plus an allow-list for the numpy scalars
        the reference stores in ``best_result``; only if that fails (and ``allow_pickle`` is left on -- checkpoints
        from untrusted sources should be loaded with ``allow_pickle=False``) with the unrestricted unpickler the
        reference's ``torch.load`` used."""
        try:
            safe = [np.dtype, np.float64, np.float32, np.int64]
            core = getattr(np, "_core", None) or getattr(np, "core")
            safe += [core.multiarray.scalar, core.multiarray._reconstruct, np.ndarray]
            safe += [type(np.dtype(t)) for t in (np.float64, np.float32, np.int64)]
            with torch.serialization.safe_globals(safe):
                ckpt = torch.load(file_path, map_location=self.device, weights_only=True)
        except Exception:
            if not allow_pickle:
                raise
            ckpt = torch.load(file_path, map_location=self.device, weights_only=False)
        self.model.load_state_dict(ckpt['model_state_dict'])
        self.optimizer.load_state_dict(ckpt['optimizer_state_dict'])
        self.current_epoch = ckpt['current_epoch'] + 1
        self.train_indices, self.valid_indices = ckpt['train_indices'], ckpt['valid_indices']
        self.best_result = ckpt['best_result']
        if 'amp_state_dict' in ckpt:
            self.amp_state_dict = ckpt['amp_state_dict']
        if 'scheduler_state_dict' in ckpt and self.scheduler is not None:
            self.scheduler.load_state_dict(ckpt['scheduler_state_dict'])
