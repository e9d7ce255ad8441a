"""Where does an end-to-end training step spend its time?  Host wall-clock marks around the prefetcher / step / item()."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet3d_b200
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = unet3d_b200.ResUnet3D(out_channels=3).to(dev).train()
loss_fn = unet3d_b200.DiceLoss()
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
h_img = torch.randn(2, 1, 128, 128, 128).pin_memory()
h_lab = torch.randint(0, 3, (2, 128, 128, 128)).pin_memory()
stepper = unet3d_b200.GraphedTrainStep(model, loss_fn, opt, warmup=3)
for _ in range(6):
    stepper(h_img.to(dev), h_lab.to(dev))
torch.cuda.synchronize()
mode = sys.argv[1] if len(sys.argv) > 1 else "prefetch"
for rep in range(2):
    t = [time.perf_counter()]
    if mode == "prefetch":
        for b in unet3d_b200.DevicePrefetcher([{"image": h_img, "label": h_lab} for _ in range(5)], dev):
            t.append(time.perf_counter())
            loss = stepper(b["image"], b["label"])[0]
            t.append(time.perf_counter())
            loss.item()
            t.append(time.perf_counter())
    else:
        for _ in range(5):
            img, lab = h_img.to(dev, non_blocking=True), h_lab.to(dev, non_blocking=True)
            t.append(time.perf_counter())
            loss = stepper(img, lab)[0]
            t.append(time.perf_counter())
            loss.item()
            t.append(time.perf_counter())
    print(mode, "total ms", round((t[-1] - t[0]) * 1e3, 2), "marks (got batch, enqueued step, item) ms:",
          [round((b - a) * 1e3, 2) for a, b in zip(t[:-1], t[1:])], flush=True)
