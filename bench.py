#!/usr/bin/env python
"""Benchmark of the 3D U-Net training hot path (BASELINE.json metric: train voxels/s @128^3 patch).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU

A "step" is one full training step of ResUnet3D(num_pool=4, num_features=30, out_channels=3) on a batch of
2 x 1 x 128^3 synthetic patches per GPU (BASELINE.json configs[1]): zero_grad, forward, Dice loss, backward,
[NCCL gradient all-reduce when N > 1], Adam step.  One JSON line is printed by rank 0 (see the contract in
the task description): `value` times the step with the batch resident in HBM, `e2e` times the same step
through the public API from pinned host memory (H2D of image + label, D2H of the loss) each step.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PATCH = (128, 128, 128)
BATCH = 2
FWD_BWD_FLOP_PER_VOXEL = 1885512          # BASELINE.md section 3 (default net, out=3): conv FLOPs fwd + bwd
WORKLOAD = "cfg-2: ResUnet3D(4,30,out=3) train step, batch 2 x 1x128^3 per GPU, DiceLoss, Adam"


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured"
    except Exception:
        return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def _reference_or_port():
    """The reference's own modules from oracle/_ref (copied there by oracle/build_ref.py at build time; kind "reference")
    or, when that directory did not travel, the oracle port (kind "port")."""
    from oracle import build_ref
    mods = build_ref.load()
    return (mods, "reference") if mods is not None else (None, "port")


class CpuStepper:
    """One training step of the reference algorithm on the host cores, the SAME timed region as the CUDA arm:
    zero_grad + forward (train mode, Dropout3d active) + DiceLoss + backward + Adam step, fp32, every core."""

    def __init__(self, shape):
        import torch
        torch.set_num_threads(os.cpu_count() or 1)             # BASELINE.md 4.2 -- also under torchrun (OMP_NUM_THREADS=1)
        mods, self.kind = _reference_or_port()
        torch.manual_seed(0)
        self.torch = torch
        n, d, h, w = shape
        self.x = torch.randn(n, 1, d, h, w, generator=torch.Generator().manual_seed(1234))
        self.y = torch.randint(0, 3, (n, d, h, w), generator=torch.Generator().manual_seed(4321))
        self.voxels = n * d * h * w
        if mods is not None:
            network, loss = mods
            self.model = network.ResUnet3D(num_pool=4, num_features=30, in_channels=1, out_channels=3).train()
            self.loss = loss.DiceLoss()
            self.params = list(self.model.parameters())
        else:
            from oracle import unet3d_oracle as O
            import unet3d_b200
            self.O = O
            self.sd = {k: v.detach().clone().requires_grad_(True)
                       for k, v in unet3d_b200.ResUnet3D(out_channels=3).state_dict().items()}
            self.params = list(self.sd.values())
        self.opt = torch.optim.Adam(self.params, lr=1e-4)
        self.cores = torch.get_num_threads()

    def step(self):
        t0 = time.perf_counter()
        self.opt.zero_grad(set_to_none=True)
        if self.kind == "reference":
            loss = self.loss(self.model(self.x), self.y)
        else:
            loss = self.O.dice_loss(self.O.resunet3d_forward(self.sd, self.x, masks=self.O.DropoutMasks(train=True)), self.y)
        loss.backward()
        self.opt.step()
        return time.perf_counter() - t0


def _host_ram_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:
        return 0.0


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle/_ref when present) on the box's host
    cores, --steps / --warmup honoured.  The step is the metric's own config (2 x 1x128^3) when one such step fits the
    time budget (about 200 s for the whole run) and 64 GB of free host RAM, else a bounded sample (1 x 1x64^3: 1/16 of
    the voxels; the network's FLOPs and activation bytes are exactly linear in voxels) -- `cpu_baseline.sample` says which."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    budget_s = float(os.environ.get("U3D_REF_BUDGET_S", "200"))
    small = CpuStepper((1, 64, 64, 64))
    t_small = min(small.step() for _ in range(2))
    full_cfg = None
    est_full = t_small * 16.0 * 1.1
    if est_full * (steps + warmup + 1) <= budget_s and _host_ram_gb() >= 64.0 and not args.no_full_step:
        try:
            full_cfg = CpuStepper((BATCH, *PATCH))
        except Exception:
            full_cfg = None
    runner, sample = (full_cfg, f"{steps} steps of the full config, {BATCH} x 1x128^3 per step") if full_cfg is not None else \
                     (small, f"{steps} steps of 1 x 1x64^3 (1/16 of the 2x128^3 step's voxels; FLOPs and bytes linear in voxels)")
    for _ in range(warmup):
        runner.step()
    times = [runner.step() for _ in range(steps)]
    sec = sum(times) / len(times)
    value = runner.voxels / sec
    one_full = None
    if full_cfg is None and not args.no_full_step and _host_ram_gb() >= 64.0 and est_full <= 120.0:
        # one real step of the metric's config, as a check of the linear extrapolation
        try:
            one = CpuStepper((BATCH, *PATCH))
            dt = one.step()
            one_full = {"ms": dt * 1e3, "voxels_per_s": one.voxels / dt, "note": "ONE cold step of 2 x 1x128^3 (no warm-up)"}
        except Exception as e:      # pragma: no cover
            one_full = {"error": repr(e)[:200]}
    line = {"impl": "reference", "metric": "train voxels/s (fwd+bwd)", "value": value, "unit": "voxels/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": BATCH, "patch": list(PATCH),
                       "timed_region": "zero_grad + forward + DiceLoss + backward + Adam step",
                       "note": ("the reference's own network.py / loss.py (oracle/_ref), " if runner.kind == "reference" else
                                "oracle port of network.py / loss.py, ") + "torch CPU fp32, train mode, all host cores"},
            "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": runner.cores, "kind": runner.kind, "sample": sample,
                             "host_cpus": os.cpu_count(), "full_config_step": one_full},
            "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cudnn_bar(dev, l2_flush, steps=3):
    """SURVEY.md 2.2 / 8d "secondary bar": the UNMODIFIED algorithm (the reference's network.py + loss.py from
    oracle/_ref, else the oracle port) through PyTorch / cuDNN on this GPU at cfg-2 -- fp32 as PyTorch runs it by default
    (cuDNN may use TF32) and bf16 autocast with channels_last_3d.  Context for `value`, not credit: it answers whether
    the hand-written path beats the library path the reference would dispatch to on the same box."""
    import torch
    mods, kind = _reference_or_port()
    out = {"impl": kind, "config": "cfg-2: 2 x 1x128^3, train mode, zero_grad + forward + DiceLoss + backward + Adam(fused)"}
    torch.manual_seed(0)
    x = torch.randn(BATCH, 1, *PATCH, device=dev)
    y = torch.randint(0, 3, (BATCH, *PATCH), device=dev)
    for name in ("fp32", "bf16_autocast_channels_last_3d"):
        try:
            if mods is not None:
                net = mods[0].ResUnet3D(num_pool=4, num_features=30, in_channels=1, out_channels=3).to(dev).train()
                loss_fn = mods[1].DiceLoss()
                fwd = lambda inp: loss_fn(net(inp), y)
                params = list(net.parameters())
            else:
                from oracle import unet3d_oracle as O
                import unet3d_b200
                sd = {k: v.detach().to(dev).requires_grad_(True)
                      for k, v in unet3d_b200.ResUnet3D(out_channels=3).state_dict().items()}
                fwd = lambda inp: O.dice_loss(O.resunet3d_forward(sd, inp, masks=O.DropoutMasks(train=True)), y.cpu()).to(dev)
                params = list(sd.values())
                net = None
            opt = torch.optim.Adam(params, lr=1e-4, fused=True)
            inp = x
            if name != "fp32":
                inp = x.contiguous(memory_format=torch.channels_last_3d)
                if net is not None:
                    net = net.to(memory_format=torch.channels_last_3d)

            def step():
                opt.zero_grad(set_to_none=True)
                if name == "fp32":
                    loss = fwd(inp)
                else:
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        loss = fwd(inp)
                loss.backward()
                opt.step()
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            evs = []
            for _ in range(steps):
                l2_flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step()
                e1.record()
                evs.append((e0, e1))
            torch.cuda.synchronize()
            ms = sum(a.elapsed_time(b) for a, b in evs) / steps
            out[name] = {"ms_per_step": ms, "voxels_per_s": BATCH * PATCH[0] ** 3 / (ms * 1e-3),
                         "peak_mem_gib": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1)}
            del opt, params, net
        except Exception as e:                      # pragma: no cover - context only, never fails the bench
            out[name] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
    out["note"] = ("the reference returns its loss as a CPU tensor built from per-class .item()-like copies (loss.py:112-118), so "
                   "each step of this arm contains C host synchronisations, as the reference's own training loop does")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-step", action="store_true", help="reference arm: never run a 2 x 128^3 step on the CPU")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the context measurements (other precision, cfg-3, cfg-5, cuDNN bar)")
    ap.add_argument("--patch", type=int, default=PATCH[0], help="cubic patch edge (default 128 = the metric's config)")
    ap.add_argument("--no-infer", action="store_true", help="skip the sliding-window inference measurement (cfg-4)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--eager", action="store_true",
                    help="launch every kernel of the step from Python instead of replaying the captured step "
                         "(GraphedTrainStep, the library's Trainer(cuda_graph=True) path)")
    ap.add_argument("--two-graph", action="store_true",
                    help="N > 1: round-1 behaviour -- two graphs per step with an eager NCCL all-reduce between them "
                         "(default: ONE graph per step with the all-reduce captured inside it)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="N > 1: do not send the first ~85 %% of the gradient bytes during the encoder half of the backward pass")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import unet3d_b200
    from unet3d_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the all-reduce of the first gradient chunks runs UNDER the backward pass and takes SMs from the persistent
        # one-CTA-per-SM tensor kernels: 16-32 CTAs measured best (2 B200s: 32 / 16 / 8 / 4 CTAs -> 19.57 / 19.65 / 20.03 /
        # 20.60 ms; 8 B200s: 19.79 with 16, 19.82 with 32); overridable from the environment
        os.environ.setdefault("NCCL_MAX_CTAS", "16")
        # NCCL_DEBUG=VERSION / WARN writes a version banner to NCCL's log file, stdout by default: rank 0's stdout must
        # carry ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":      # this image's default: a bare printf to stdout
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = args.steps
    pe = args.patch
    ops.REAL_CHANNELS = {32: 30, 64: 60, 128: 120}      # HBM-kernel bytes are counted on the net's unpadded widths
    torch.manual_seed(0)
    model = unet3d_b200.ResUnet3D(num_pool=4, num_features=30, in_channels=1, out_channels=3).to(dev).train()
    model.precision = args.precision
    loss_fn = unet3d_b200.DiceLoss()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)

    g = torch.Generator().manual_seed(1234 + rank)
    h_img = torch.randn(BATCH, 1, pe, pe, pe, generator=g).pin_memory()
    h_lab = torch.randint(0, 3, (BATCH, pe, pe, pe), generator=torch.Generator().manual_seed(4321 + rank)).pin_memory()
    d_img, d_lab = h_img.to(dev), h_lab.to(dev)
    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def allreduce_grads():
        unet3d_b200.parallel.all_reduce_gradients(model)      # bucketed NCCL all-reduce (mean); no-op for N = 1

    def eager_step(img, lab):
        opt.zero_grad(set_to_none=True)
        out = model(img)
        loss = loss_fn(out, lab)
        loss.backward()
        allreduce_grads()
        opt.step()
        return loss

    # default: the whole step replayed as a CUDA graph (one graph at N = 1; at N > 1 two graphs with the NCCL all-reduce
    # issued eagerly between them) -- the same kernels, without ~530 Python-side launches per step
    stepper = None if args.eager else unet3d_b200.GraphedTrainStep(model, loss_fn, opt, warmup=3,
                                                                   capture_collectives=not args.two_graph,
                                                                   overlap=not args.no_overlap)

    def step(img, lab):
        if stepper is None:
            return eager_step(img, lab)
        return stepper(img, lab)[0]

    def timed_e2e(n_steps):
        """End to end through the public API: pinned host batches -> DevicePrefetcher (the upload of batch i+1 runs on a
        side stream under the step of batch i; every batch is uploaded inside the timed region) -> training step ->
        loss.item().  One pair of events around the whole loop, first upload included."""
        l2_flush.zero_()
        host_batches = [{"image": h_img, "label": h_lab} for _ in range(n_steps)]
        staging = timed_e2e.__dict__.setdefault("staging", {})        # one set of staging buffers for all e2e loops
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        marks = [time.perf_counter()]
        for b in unet3d_b200.DevicePrefetcher(host_batches, dev, staging):
            loss = step(b["image"], b["label"])
            _ = loss.item()                                   # D2H of the step's result
            marks.append(time.perf_counter())
        e1.record()
        torch.cuda.synchronize()
        if os.environ.get("U3D_BENCH_DEBUG"):
            print(f"[bench rank {rank}] e2e host ms per step: {[round((b - a) * 1e3, 2) for a, b in zip(marks[:-1], marks[1:])]}, "
                  f"events {e0.elapsed_time(e1):.2f} ms, graph {'yes' if stepper is not None and stepper.graph is not None else 'no'}",
                  file=sys.stderr, flush=True)
        return e0.elapsed_time(e1) / n_steps

    def timed(n_steps, e2e, lead_sleep=False):
        if e2e:
            return timed_e2e(n_steps)
        evs = []
        for _ in range(n_steps):
            l2_flush.zero_()                                  # flush L2 between timed iterations
            if lead_sleep:
                # per-kernel profiling pass (eager launches, an event pair around each): a ~40 ms spin kernel lets the
                # host enqueue the whole step ahead of the GPU, so that every event interval is kernel time and not the
                # host's launch latency (which used to inflate every launch shorter than ~10 us)
                torch.cuda._sleep(int(0.040 * 1.9e9))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if e2e:
                img = h_img.to(dev, non_blocking=True)
                lab = h_lab.to(dev, non_blocking=True)
                loss = step(img, lab)
                _ = loss.item()                               # D2H of the step's result
            else:
                step(d_img, d_lab)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        if os.environ.get("U3D_BENCH_DEBUG"):
            print(f"[bench rank {rank}] e2e={e2e} per-step ms: {[round(a.elapsed_time(b), 2) for a, b in evs]}", file=sys.stderr, flush=True)
        return sum(a.elapsed_time(b) for a, b in evs) / n_steps

    for i in range(max(W, 3) + (0 if stepper is None else 2)):      # 3 eager steps, then capture + one replay
        step(d_img, d_lab)
        if i == 0 and world > 1 and stepper is None and not args.no_overlap:
            # eager launches: engines exist after the first forward; from now on 1 / world is folded into the gradient
            # unpack and the first ~85 % of the gradient bytes are all-reduced during the encoder half of the backward pass
            unet3d_b200.parallel.prescale_gradients(model)
            unet3d_b200.parallel.overlap_gradient_all_reduce(model)
    torch.cuda.synchronize()
    ops.check_device_errors()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ops.LAUNCHES = 0
    ms = timed(K, e2e=False)
    launches = ops.LAUNCHES if stepper is None else stepper.launches_per_replay * K
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # per-kernel CUDA events for the roofline: a REPEAT of the timed region (same steps, same inputs, same clocks window)
    # with an event pair around every launch -- kept out of `value` because ~700 event records per step cost ~1.5 ms
    used_graph = stepper is not None
    ops.PROFILE = []
    graphed, stepper = stepper, None          # the event pairs need eager launches
    timed(K, e2e=False, lead_sleep=True)
    stepper = graphed
    dp_mode = graphed.mode if graphed is not None else "eager"
    prof = ops.PROFILE
    ops.PROFILE = None
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    timed(1, e2e=True)                                   # untimed: first use of the pinned-memory path
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()                                    # rank 0 just spent 150 ms stopping the clock sampler
    ms_e2e = timed(max(2, min(K, 5)), e2e=True)
    ops.check_device_errors()

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    vox = BATCH * pe ** 3 * world

    # dominant kernel: conv_gemm (forward + data-gradient launches) -- algorithmic FLOPs / CUDA-event time
    roof = None
    if rank == 0 and prof:
        by, mem, big = {}, {}, {}
        for name, flops, a, b, nbytes in prof:
            d = (by if flops > 0 else mem).setdefault(name, [0.0, 0.0, 0])
            d[0] += flops if flops > 0 else nbytes
            d[1] += a.elapsed_time(b)
            d[2] += 1
            if flops <= 0:
                big.setdefault(name, []).append((nbytes, a.elapsed_time(b)))
        # the same kernels on their LARGEST tensors only (level 0: 2 x 128^3 x 30 channels): the per-step mean above also
        # holds the 1-16 MB tensors of levels 2-4, which are launch-latency-bound whatever the kernel does
        largest = {}
        for name, rows in big.items():
            top = max(r[0] for r in rows)
            sel = [r for r in rows if r[0] >= 0.9 * top]
            tb, tt = sum(r[0] for r in sel), sum(r[1] for r in sel)
            largest[name] = (len(sel) // K, tb / (tt * 1e-3) / 1e9 if tt > 0 else None)
        peak, hbm, src = read_peaks()
        name = max(by, key=lambda k: by[k][1])
        fl, tms, cnt = by[name]
        ach = fl / (tms * 1e-3) / 1e12
        # traffic: dram__bytes_read + dram__bytes_write of ONE launch from the committed ncu --set full capture
        # (profiles/r02_ncu_kernels.txt: the level-0 30->30 forward launch, 203.8 algorithmic GFLOP, 537 MB algorithmic
        # bytes = 16-bit input + output; 268.6 MB read + 220.2 MB written); the per-step `achieved` above is the
        # FLOP-weighted mean over all 92 launches of the step (level 0: ~800 TFLOP/s, strided / transposed / 8^3-grid
        # layers far below: profiles/r02_step_table.txt)
        roof = {"bound": "tensor", "kernel": name, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": 488.8e6,
                "traffic_note": "level-0 conv 30->30 k3 forward launch (203.8 GFLOP, 537e6 algorithmic bytes), ncu --set full "
                                "capture of round 2: profiles/r02_ncu_kernels.txt (a bench run cannot measure DRAM bytes itself)",
                "peak_source": f"{src} (bf16_tflops_sustained)",
                "launches_per_step": cnt // K, "kernel_ms_per_step": tms / K,
                "per_kernel_ms_per_step": {k: v[1] / K for k, v in by.items()},
                "per_kernel_tflops": {k: (v[0] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else None) for k, v in by.items()},
                # the HBM-bound kernels of the step: algorithmic bytes / CUDA-event time against the measured copy bandwidth
                "hbm_peak_gbs": hbm,
                "hbm_kernels": {k: {"launches_per_step": v[2] // K, "ms_per_step": round(v[1] / K, 4),
                                    "gbs": round(v[0] / (v[1] * 1e-3) / 1e9, 1) if v[1] > 0 else None,
                                    "frac": round(v[0] / (v[1] * 1e-3) / 1e9 / hbm, 3) if v[1] > 0 else None,
                                    "largest_tensor_launches_per_step": largest[k][0],
                                    "largest_tensor_gbs": round(largest[k][1], 1) if largest[k][1] else None,
                                    "largest_tensor_frac": round(largest[k][1] / hbm, 3) if largest[k][1] else None}
                                for k, v in sorted(mem.items(), key=lambda kv: -kv[1][1])}}

    # ---- context measurements (extra keys; not part of `value`): the other 16-bit storage format, BASELINE.json's cfg-3
    # and cfg-5 at full size, and the same algorithm through PyTorch / cuDNN on this GPU (SURVEY.md 8d "secondary bar")
    extras = {}
    if not args.no_extras and pe == 128:
        stepper = graphed = None                # frees the captured step's memory pool
        opt.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()

        def time_config(make_model, shape, precision, steps=5):
            """ms per full training step (graph replay, L2 flushed between steps, max over ranks) of another config."""
            torch.manual_seed(0)
            m = make_model().to(dev).train()
            m.precision = precision
            o = torch.optim.Adam(m.parameters(), lr=1e-4, fused=True)
            gx = torch.Generator(device=dev).manual_seed(99 + rank)
            x = torch.randn(*shape, device=dev, generator=gx)
            y = torch.randint(0, 3, (shape[0], *shape[2:]), device=dev, generator=gx)
            st = unet3d_b200.GraphedTrainStep(m, unet3d_b200.DiceLoss(), o, warmup=2)
            for _ in range(4):                      # 2 eager steps, capture + replay, one more replay
                st(x, y)
            torch.cuda.synchronize()
            ops.check_device_errors()
            evs = []
            for _ in range(steps):
                l2_flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                st(x, y)
                e1.record()
                evs.append((e0, e1))
            torch.cuda.synchronize()
            t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs) / steps], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            peak = torch.cuda.max_memory_allocated(dev) / 2 ** 30
            del st, m, o, x, y
            torch.cuda.empty_cache()
            return float(t.item()), peak

        def entry(ms_, voxels, flop_per_voxel, peak_gib, **kw):
            return {"ms_per_step": ms_, "voxels_per_s": voxels / (ms_ * 1e-3), "algorithmic_tflops_per_gpu":
                    flop_per_voxel * voxels / world / (ms_ * 1e-3) / 1e12, "peak_mem_gib": round(peak_gib, 1), **kw}

        default_net = lambda: unet3d_b200.ResUnet3D(num_pool=4, num_features=30, in_channels=1, out_channels=3)
        other = "fp16" if args.precision == "bf16" else "bf16"
        ms_o, pk = time_config(default_net, (BATCH, 1, pe, pe, pe), other)
        extras[other] = entry(ms_o, BATCH * pe ** 3 * world, FWD_BWD_FLOP_PER_VOXEL, pk,
                              note=f"the same cfg-2 step with {other} storage (same tcgen05 kind::f16 rate)")
        # cfg-3: KiTS19-shaped 160x160x80 patches, global batch 16 over 2 / 4 / 8 GPUs (2 per GPU when run on one GPU)
        b3 = 16 // world if world in (2, 4, 8) else 2
        ms_3, pk = time_config(default_net, (b3, 1, 160, 160, 80), args.precision)
        extras["cfg3"] = entry(ms_3, b3 * 160 * 160 * 80 * world, FWD_BWD_FLOP_PER_VOXEL, pk, per_gpu_batch=b3,
                               global_batch=b3 * world, patch=[160, 160, 80])
        # cfg-5: wide / deep variant, ResUnet3D(num_pool=5, num_features=32) on one 192^3 patch per GPU
        ms_5, pk = time_config(lambda: unet3d_b200.ResUnet3D(num_pool=5, num_features=32, in_channels=1, out_channels=3),
                               (1, 1, 192, 192, 192), args.precision, steps=3)
        extras["cfg5"] = entry(ms_5, 192 ** 3 * world, 2239200, pk, per_gpu_batch=1, patch=[192, 192, 192],
                               net="ResUnet3D(num_pool=5, num_features=32, out=3), 464 M parameters")
        if rank == 0:
            extras["cudnn_bar"] = cudnn_bar(dev, l2_flush)

    # ---- second half of BASELINE.json's metric: sliding-window inference, CT volumes/s (cfg-4: 512x512x256 volume,
    # 128^3 windows at 50 % overlap = 147 windows on the reference's grid; windows are dealt to the ranks)
    infer = None
    if not args.no_infer and pe == 128:
        import numpy as np
        d_img = d_lab = None
        opt.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()
        vol = np.random.RandomState(7).standard_normal((512, 512, 256, 1)).astype(np.float32)
        # warm-up on EVERY rank: a sub-volume with >= 5 windows per rank at 8 ranks, so that the window-batch plans are
        # built (~seconds of host work on first use of a shape) and the window forward is captured before the timed call
        unet3d_b200.predict_per_patch(vol[:384, :256, :256], model, 3, (128, 128, 128), 2, verbose=False)
        runs = []
        for _ in range(3):                                  # median of three whole-volume calls
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            labels = unet3d_b200.predict_per_patch(vol, model, 3, (128, 128, 128), 2, verbose=False)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            tt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            runs.append(float(tt.item()))
        dt = sorted(runs)[1]
        n_win = len(unet3d_b200.tile_origins((512, 512, 256), (128, 128, 128), 2))
        infer = {"metric": "infer CT volumes/s", "value": 1.0 / dt, "unit": "volumes/s", "seconds_per_volume": dt, "seconds_all_runs": [round(v, 4) for v in runs],
                 "volume": [512, 512, 256], "window": [128, 128, 128], "windows": n_win, "grid": "reference (trainer.py:29-40)",
                 "blend": "uniform", "windows_per_forward": 4, "includes": "H2D of the fp32 volume from pageable host memory (x-slabs, overlapped with the window forwards), all "
                 "window forwards, blend, normalise + argmax, D2H of the uint8 label map into pinned memory" + (", all-reduce of the blend buffers" if world > 1 else ""),
                 "label_hist": [int(v) for v in np.bincount(labels.reshape(-1), minlength=3)[:3]],
                 "label_sha256": hashlib.sha256(np.ascontiguousarray(labels).tobytes()).hexdigest()[:16]}
        if world > 1:
            # the sharded label map must be the single-GPU label map, bit for bit: rank 0 recomputes the volume alone
            if rank == 0:
                single = unet3d_b200.predict_per_patch(vol, model, 3, (128, 128, 128), 2, verbose=False, distributed=False)
                infer["single_gpu_label_sha256"] = hashlib.sha256(np.ascontiguousarray(single).tobytes()).hexdigest()[:16]
                infer["labels_equal_single_gpu"] = bool(np.array_equal(single, labels))
                infer["label_mismatches_vs_single_gpu"] = int((single != labels).sum())
            dist.barrier()

        if rank == 0:
            # the steps either side of the window loop (trainer.predict_case): zoom + clip + z-score of a raw
            # 512x512x128 case onto the 512x512x256 grid, and the label map zoomed back -- HBM-bound kernels
            from unet3d_b200 import transform as T
            _, hbm, _ = read_peaks()
            raw = torch.randn(512, 512, 128, 1, device=dev)
            lab_dev = torch.from_numpy(labels).to(dev)
            table = T.normalize_table({"mean": 0.1, "std": 0.9, "pct_00_5": -2.0, "pct_99_5": 2.0})

            def kernel_ms(fn, reps=5):
                # CUDA events recorded immediately around the C-ABI call (ops._Timed), not around the Python wrapper
                fn()
                ts = []
                for _ in range(reps):
                    l2_flush.zero_()
                    ops.PROFILE = []
                    fn()
                    torch.cuda.synchronize()
                    ts.append(sum(a.elapsed_time(b) for _, _, a, b, _ in ops.PROFILE))
                    ops.PROFILE = None
                return sorted(ts)[len(ts) // 2]
            up = torch.empty(512, 512, 256, 1, device=dev)
            down = torch.empty(512, 512, 128, dtype=torch.uint8, device=dev)
            ms_up = kernel_ms(lambda: T.rescale_device(raw, (1.0, 1.0, 2.0), multi_class=True, out=up, norm=table))
            ms_dn = kernel_ms(lambda: T.rescale_device(lab_dev, (1.0, 1.0, 0.5), is_label=True, num_classes=3, out=down))
            b_up, b_dn = raw.numel() * 4 + up.numel() * 4, lab_dev.numel() + down.numel()
            infer["prepost"] = {
                "zoom_linear_norm": {"in": [512, 512, 128], "out": [512, 512, 256], "ms": round(ms_up, 4),
                                     "gbs": round(b_up / ms_up / 1e6, 1), "frac_hbm": round(b_up / ms_up / 1e6 / hbm, 3)},
                "zoom_label": {"in": [512, 512, 256], "out": [512, 512, 128], "ms": round(ms_dn, 4),
                               "gbs": round(b_dn / ms_dn / 1e6, 1), "frac_hbm": round(b_dn / ms_dn / 1e6 / hbm, 3)},
                "note": "algorithmic bytes (volume in + volume out) / CUDA-event time incl. the per-axis table kernel; "
                        "bit-exact with scipy.ndimage.zoom(order=1): float64 corner sums"}
            del raw, lab_dev, up, down

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        c = CpuStepper((1, 64, 64, 64))
        c.step()
        sec = sum(c.step() for _ in range(4)) / 4
        cpu = {"value": c.voxels / sec, "unit": "voxels/s", "cores": c.cores, "kind": c.kind, "host_cpus": os.cpu_count(),
               "sample": "4 steps (after 1 warm-up) of 1 x 1x64^3 = 1/16 of the 2x128^3 step: zero_grad + forward + DiceLoss + "
                         "backward + Adam, fp32, train mode, " + ("the reference's own network.py / loss.py (oracle/_ref)"
                                                                  if c.kind == "reference" else "oracle port of network.py / loss.py")}

    if rank == 0:
        line = {"metric": "train voxels/s (fwd+bwd)", "value": vox / (ms * 1e-3), "unit": "voxels/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": {"workload": WORKLOAD if pe == 128 else f"REDUCED patch {pe}^3 (not the metric's config)",
                           "global_batch": BATCH * world, "patch": [pe, pe, pe], "parallelism": f"dp{world}", "cuda_graph": used_graph, "dp_mode": dp_mode if world > 1 else None,
                           "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS") if world > 1 else None,
                           "sm_limit_window_launches": list(unet3d_b200.engine.NCCL_WINDOW) if world > 1 else None,
                           "timed_region": "zero_grad + forward + DiceLoss + backward + grad all-reduce (N>1) + Adam step",
                           "l2": "256 MB buffer written between timed iterations (L2 flush); activations per step >> L2",
                           "tensor_frac_of_step": (FWD_BWD_FLOP_PER_VOXEL * BATCH * pe ** 3 / (ms * 1e-3) / 1e12) /
                                                  read_peaks()[0]},
                "e2e": {"value": vox / (ms_e2e * 1e-3), "unit": "voxels/s",
                        "h2d_bytes_per_step": (h_img.numel() * 4 + h_lab.numel() * 8) * world, "d2h_bytes_per_step": 4 * world,
                        "ms_per_step": ms_e2e},
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "infer": infer,
                **extras}
        print(json.dumps(line), flush=True)
    if world > 1:
        # tear-down: graphs that hold captured NCCL kernels must go before the process group does, and a communicator
        # destroy that blocks (seen with captured collectives) must never keep the job alive after the result is printed
        stepper = graphed = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        killer = threading.Timer(20.0, lambda: os._exit(0))
        killer.daemon = True
        killer.start()
        dist.destroy_process_group()
        killer.cancel()


if __name__ == "__main__":
    main()
