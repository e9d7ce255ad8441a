"""InstanceNorm vs BatchNorm versions of the same 4-channel attention net against the fp32 oracle (test infrastructure:
the error levels quoted in tests/test_variants_gpu.py).  python tests/tools/bn_vs_in_check.py  (one B200)"""
import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
import unet3d_b200
from oracle import unet3d_oracle as O
def build(bn):
    nk = {'norm_op': torch.nn.BatchNorm3d} if bn else {}
    nd = {**nk, 'dropout_op': None}
    return unet3d_b200.Unet(1, 3, unet3d_b200.generate_paired_features(2, 4), pool_block=unet3d_b200.ResBlock,
                            pool_kwargs={'stride': 2, **nd}, up_kwargs={'attention': True, **nk},
                            encode_block=unet3d_b200.ResBlockStack, encode_kwargs=nd,
                            encode_kwargs_fn=lambda level: {'num_stacks': max(level, 1)},
                            decode_block=unet3d_b200.ResBlock, decode_kwargs=nd)
rel=lambda a,b: ((a-b).norm()/b.norm().clamp_min(1e-20)).item()
for bn in (False, True):
  for prec in ("fp16","bf16"):
    torch.manual_seed(3)
    m=build(bn)
    sd={"net."+k: v.clone() for k,v in m.state_dict().items()}
    x=torch.randn(2,1,16,16,16); y=torch.randint(0,3,(2,16,16,16))
    m=m.cuda().train(); m.precision=prec
    lg=m(x.cuda()); loss=unet3d_b200.HybirdLoss(weight_v=[1,148,191],alpha=0.9,beta=0.1)(lg,y.cuda()); loss.backward()
    for k,v in sd.items():
        if v.dtype.is_floating_point and "running" not in k: v.requires_grad_(True)
    rl=O.resunet3d_forward(sd,x,2,4,attention=True,bn_train=True)
    O.hybrid_loss(rl,y,weight_v=[1,148,191],alpha=0.9,beta=0.1).backward()
    rs=[]
    for n,p in m.named_parameters():
        rg=sd["net."+n].grad
        if rg is None or p.grad is None: continue
        if n.endswith("bias") and ("conv1." in n or "conv2." in n or "up.0" in n): continue
        rs.append(rel(p.grad.cpu(),rg))
    print(f"bn={bn} {prec}: logits rel {rel(lg.detach().cpu(), rl.detach()):.3e}; grad rel median {np.median(rs):.3e} max {max(rs):.3e}")
