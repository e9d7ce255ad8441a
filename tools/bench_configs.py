"""Full-size timing of BASELINE.json's other single-GPU configurations (parity for them runs at reduced sizes in
tests/test_model_gpu.py): cfg-3's per-GPU share (2 x 160x160x80, default net) and cfg-5 (ResUnet3D(5, 32) on 192^3).
Usage: python tools/bench_configs.py [cfg3|cfg5|all]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet3d_b200

dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
cases = {"cfg3": (dict(num_pool=4, num_features=30, out_channels=3), (2, 1, 160, 160, 80), 1885512),
         "cfg5": (dict(num_pool=5, num_features=32, out_channels=3), (1, 1, 192, 192, 192), 2239200)}
for name, (kw, shape, flop_per_voxel) in cases.items():
    if which not in ("all", name):
        continue
    torch.manual_seed(0)
    model = unet3d_b200.ResUnet3D(**kw).to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    loss_fn = unet3d_b200.DiceLoss()
    x = torch.randn(*shape, device=dev)
    y = torch.randint(0, 3, (shape[0], *shape[2:]), device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(model(x), y)
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        l = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        l = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    vox = shape[0] * shape[2] * shape[3] * shape[4]
    print(f"{name}: {kw} batch {shape}: {ms:.2f} ms/step, {vox / ms / 1e3:.1f} M voxels/s, "
          f"{flop_per_voxel * vox / ms / 1e9:.0f} TFLOP/s algorithmic, loss {l.item():.4f}, "
          f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del model, opt, x, y
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
