"""Three CAM forward calls and one training step at BASELINE config 2 (b=256, nc=5, D=512): the
target of the per-kernel ncu launch list of the CAM (scripts/gpu_profile_round.sh)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtc_b200.model import PretrainedCLIP_finaltf, clip_loss  # noqa: E402
from vtc_b200.synthetic import make_batch_pair, make_cam_inputs  # noqa: E402

dev = torch.device("cuda:0")
b, D, nc = 256, 512, 5
vis, txt = make_batch_pair(b, D, seed=1023)
main, aux = make_cam_inputs(b, nc, D, seed=1023)
model = PretrainedCLIP_finaltf(D).to(dev)
for blk in model.final_transformer.resblocks:
    torch.nn.init.normal_(blk.mlp.c_proj.weight, std=0.02)
    torch.nn.init.normal_(blk.attn.out_proj.weight, std=0.02)
model.eval()
with torch.no_grad():
    for _ in range(3):
        out = model._adapt_feature(main.to(dev), aux.to(dev))
model.train()
model.random_skip_adapter = False
title = txt.to(dev).requires_grad_(True)
loss = clip_loss(model(vis.to(dev), title, aux.to(dev).permute(1, 0, 2).contiguous()), {})
loss.backward()
torch.cuda.synchronize()
print(float(out.norm(dim=-1).mean()), float(loss))
