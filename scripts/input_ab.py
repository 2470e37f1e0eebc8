"""A/B of the two HWC -> CHW ToTensor kernels (row-wise vs 1-D with 64-bit divisions) on a B200:
bit-exact check of both against ``x.permute(0,3,1,2).float().div(255)``, then timing at the
workload and saturating sizes.  One JSON line per measurement."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench_roofline import to_tensor_u8
from marlclassification_b200.input_pipeline import images_u8_to_f32

dev = torch.device("cuda", 0)
SHAPES = [(8, 256, 256, 3), (2, 600, 600, 3), (32, 28, 28, 3), (3, 5, 4, 3), (1, 1, 8, 4), (5, 7, 12, 2), (2, 9, 1028, 3),
          (4, 33, 36, 4), (300, 2, 4, 3)]
for mode in ("1", "0"):
    os.environ["MARLC_U8_ROWS"] = mode  # read by the library at every call
    ok = True
    for b, h, w, c in SHAPES:
        g = torch.Generator().manual_seed(b * 100 + w)
        src = torch.randint(0, 256, (b, h, w, c), generator=g, dtype=torch.uint8)
        out = images_u8_to_f32(src.to(dev), hwc=True)
        same = torch.equal(out.cpu(), src.permute(0, 3, 1, 2).float().div(255))
        ok &= same
        if not same:
            print(json.dumps({"rows": mode, "shape": [b, h, w, c], "bit_exact": False}))
    print(json.dumps({"rows": mode, "bit_exact_all_shapes": ok, "shapes": len(SHAPES)}), flush=True)
    for args, reps in (((8, 3, 256, 256), 100), ((1365, 3, 256, 256), 20), ((248, 3, 600, 600), 20)):
        r = to_tensor_u8(*args, dev, reps=reps)
        r["rows"] = mode
        print(json.dumps(r), flush=True)
