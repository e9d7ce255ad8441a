import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import unet3d_b200
torch.manual_seed(0)
model = unet3d_b200.ResUnet3D(out_channels=3).cuda()
vol = np.random.RandomState(7).standard_normal((512, 512, 256, 1)).astype(np.float32)
for i in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    lab = unet3d_b200.predict_per_patch(vol if i else vol[:384,:256,:256], model, 3, (128,128,128), 2, verbose=False)
    torch.cuda.synchronize(); print("call", i, time.perf_counter()-t0, flush=True)
