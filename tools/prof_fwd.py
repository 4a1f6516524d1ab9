"""torch.profiler kernel table of the C1 forward (eager), to spot the non-library glue kernels."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import si_mamba_b200 as sm  # noqa: E402
from oracle import tokenizer  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
torch.manual_seed(0)
model = sm.PointMamba(sm.finetune_modelnet()).cuda().eval()
pts = tokenizer.synthetic_clouds(32, 1024, 1, "surface").cuda()


def fwd():
    with torch.no_grad():
        if prec == "bf16":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return model(pts)
        return model(pts)


for _ in range(3):
    fwd()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    for _ in range(3):
        fwd()
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.device_time_total > 0 and e.key.startswith("aten::")]
rows.sort(key=lambda e: -e.device_time_total)
for e in rows[:28]:
    print(f"{e.device_time_total / 3:9.1f} us  x{e.count // 3:3d}  {e.key:28s} {str(e.input_shapes)[:110]}")
