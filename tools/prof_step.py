"""torch.profiler kernel table of one training step of config C2 / C3 / C4 (see tools/step_bench.py)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import si_mamba_b200 as sm  # noqa: E402
from oracle import tokenizer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
dev = torch.device("cuda", 0)
torch.manual_seed(0)
if name == "C2":
    cfg = sm.finetune_scan_hardest()
    model = sm.PointMamba(cfg).to(dev).train()
    pts = tokenizer.synthetic_clouds(32, 2048, 2000, "surface").to(dev)
    label = torch.randint(0, cfg.cls_dim, (32,), device=dev)

    def step():
        model.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = model(pts)
        torch.nn.functional.cross_entropy(logits.float(), label).backward()
elif name == "C4":
    cfg = sm.part_seg_config()
    model = sm.get_model(50, cfg).to(dev).train()
    pts = tokenizer.synthetic_clouds(16, 2048, 4000, "surface").to(dev).transpose(1, 2).contiguous()
    cls = torch.nn.functional.one_hot(torch.randint(0, 16, (16,), device=dev), 16).float()
    target = torch.randint(0, 50, (16, 2048), device=dev)

    def step():
        model.zero_grad(set_to_none=True)
        out = model(pts, cls)
        torch.nn.functional.nll_loss(out.reshape(-1, 50), target.reshape(-1)).backward()
else:
    cfg = sm.pretrain()
    model = sm.Point_MAE_Mamba(cfg).to(dev).train()
    pts = tokenizer.synthetic_clouds(16, 1024, 3000, "surface").to(dev)

    def step():
        model.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = model(pts)
        loss.backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
