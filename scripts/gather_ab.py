"""A/B of the two bodies of patch_gather_kernel (multiply-high vs runtime division) on a B200:
bit-exact check of both against plain indexing on ragged shapes, then timing at the workload
and saturating sizes bench_roofline.py reports.  Prints one JSON line per measurement."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench_roofline import gather
from marlclassification_b200 import _lib

dev = torch.device("cuda", 0)
L = _lib.lib()


def reference(img, pos, f):
    na, nb = pos.shape[:2]
    out = torch.empty(na, nb, img.shape[1], f, f, device=img.device)
    for a in range(na):
        for b in range(nb):
            y, x = int(pos[a, b, 0]), int(pos[a, b, 1])
            out[a, b] = img[b, :, y:y + f, x:x + f]
    return out


SHAPES = [(3, 4, 1, 28, 28, 6), (16, 8, 3, 256, 256, 12), (5, 3, 3, 41, 37, 7), (2, 2, 3, 600, 600, 24),
          (1, 1, 2, 9, 10, 8), (4, 2, 3, 64, 48, 24), (3, 2, 1, 5, 5, 1), (2, 3, 4, 70, 66, 64), (7, 5, 3, 33, 35, 2),
          (3, 3, 8, 40, 40, 31)]
for mode in ("0", "1"):
    os.environ["MARLC_GATHER_DIV"] = mode  # read by the library at every call
    ok = True
    for na, nb, c, h, w, f in SHAPES:
        g = torch.Generator().manual_seed(na * 1000 + h)
        img = torch.rand(nb, c, h, w, generator=g).to(dev)
        pos = torch.stack([torch.randint(h - f + 1, (na, nb), generator=g),
                           torch.randint(w - f + 1, (na, nb), generator=g)], -1)
        pos[0, 0] = torch.tensor([0, 0])
        pos[-1, -1] = torch.tensor([h - f, w - f])  # the last window that still fits
        pos = pos.to(dev)
        obs = torch.full((na, nb, c, f, f), float("nan"), device=dev)
        _lib.check(L.marlc_patch_gather(img.data_ptr(), pos.data_ptr(), obs.data_ptr(), na, nb, c, h, w, f,
                                        _lib.stream_ptr(dev)))
        same = torch.equal(obs, reference(img, pos, f))
        ok &= same
        if not same:
            print(json.dumps({"div": mode, "shape": [na, nb, c, h, w, f], "bit_exact": False}))
    print(json.dumps({"div": mode, "bit_exact_all_shapes": ok, "shapes": len(SHAPES)}), flush=True)
    for args, reps in (((16, 8, 3, 256, 256, 12), 200), ((256, 256, 3, 256, 256, 12), 50), ((256, 64, 3, 600, 600, 24), 50)):
        r = gather(*args, dev, reps=reps)
        r["div"] = mode
        print(json.dumps(r), flush=True)
