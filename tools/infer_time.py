"""Wall time of predict_per_patch on the cfg-4 volume (512x512x256, 128^3 windows, 147 on the reference grid) for several
window batches; U3D_PREDICT_TIMES=1 prints the phases.  python tools/infer_time.py [window_batch ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import unet3d_b200
torch.manual_seed(0)
model = unet3d_b200.ResUnet3D(out_channels=3).cuda()
vol = np.random.RandomState(7).standard_normal((512, 512, 256, 1)).astype(np.float32)
for wb in [int(a) for a in sys.argv[1:]] or [2]:
    for i in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        lab = unet3d_b200.predict_per_patch(vol if i else vol[:384, :256, :256], model, 3, (128, 128, 128), 2, verbose=False,
                                            window_batch=wb)
        torch.cuda.synchronize(); print("window_batch", wb, "call", i, round(time.perf_counter() - t0, 4), flush=True)
