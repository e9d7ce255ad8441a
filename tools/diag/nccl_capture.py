"""Can this torch / NCCL capture an all-reduce inside a CUDA graph (sync, async + wait, side-stream overlap)?
torchrun --nproc-per-node 2 tools/diag/nccl_capture.py"""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
x = torch.ones(1 << 20, device=dev) * (rank + 1)
dist.all_reduce(x)                      # communicator exists before any capture
torch.cuda.synchronize()


def attempt(name, fn, **kw):
    g = torch.cuda.CUDAGraph()
    buf = torch.ones(1 << 20, device=dev) * (rank + 1)
    torch.cuda.synchronize()
    t0 = time.time()
    try:
        with torch.cuda.graph(g, **kw):
            out = fn(buf)
        g.replay(); g.replay()
        torch.cuda.synchronize()
        print(f"[rank {rank}] {name}: OK, value {float(out[0]):.1f} (expect {9.0 if 'twice' not in name else 9.0}) in {time.time() - t0:.2f}s", flush=True)
    except Exception as e:
        print(f"[rank {rank}] {name}: FAILED {type(e).__name__}: {str(e)[:300]}", flush=True)
        torch.cuda.synchronize()


def sync_ar(b):
    dist.all_reduce(b)
    return b

def async_ar(b):
    w = dist.all_reduce(b, async_op=True)
    c = b.new_zeros(8).add_(1)           # work on the capture stream while the collective runs
    w.wait()
    return b

attempt("sync all_reduce, default capture mode", sync_ar)
attempt("async all_reduce + wait, default capture mode", async_ar)
attempt("async all_reduce + wait, thread_local", async_ar, capture_error_mode="thread_local")
dist.barrier()
print(f"[rank {rank}] done", flush=True)
dist.destroy_process_group()
