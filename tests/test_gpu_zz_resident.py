"""``train --resident`` on the GPU: the split lives in HBM as bytes (data.ResidentImages), batches
are index_select + device ToTensor, and the run directory comes out as for the streaming loader.
(File name sorts last on purpose: the newest path runs after everything else.)"""
import json
import os

import pytest
import torch

from tests.clidata import make_image_folder

pytestmark = pytest.mark.gpu


def test_resident_batches_equal_streamed_batches_on_device(tmp_path):
    from marlclassification_b200.data import FolderDataset, ResidentImages, ShardedBatchSampler, u8_image_pipeline
    from marlclassification_b200.input_pipeline import DevicePrefetcher

    root = make_image_folder(str(tmp_path / "imgs"), classes=3, per_class=6, size=28, grey_every=2)
    ds = FolderDataset(root, u8_image_pipeline())
    sampler = ShardedBatchSampler(len(ds), 8, shuffle=True, seed=9)
    resident = ResidentImages(ds, list(range(len(ds))), sampler, "cuda")
    assert resident.images.is_cuda and resident.images.dtype == torch.uint8
    n = 0
    for staged, batch in zip(DevicePrefetcher(resident, "cuda"), sampler):
        assert staged.h2d_bytes == 0  # nothing crosses PCIe per step
        img, y = staged.deliver()
        want = torch.stack([ds[i][0] for i in batch]).permute(0, 3, 1, 2).float().div(255)
        assert torch.equal(img.cpu(), want) and y.cpu().tolist() == [ds.targets[i] for i in batch]
        n += 1
    assert n == len(sampler) == 3


def test_cli_train_resident(tmp_path):
    from marlclassification_b200.__main__ import main

    res = tmp_path / "resources"
    make_image_folder(str(res / "downloaded" / "mnist_png" / "all_png"), classes=3, per_class=17, size=28, grey_every=1)
    out = tmp_path / "run"
    main(["--run-id", "resident", "-a", "3", "--step", "4", "--cuda", "train", "--ft-extr", "mnist", "--f", "6",
          "--nb", "64", "--na", "64", "--nd", "8", "--nlb", "96", "--nla", "96", "--nb-class", "3", "--batch-size", "8",
          "--nb-epoch", "1", "--res-folder", str(res), "-o", str(out), "--workers", "2", "--resident"])
    assert os.path.exists(out / "models" / "nn_models_epoch_0.pt") and os.path.exists(out / "animated_gif.gif")
    sd = torch.load(out / "models" / "nn_models_epoch_0.pt", map_location="cpu")
    assert all(torch.isfinite(v).all() for v in sd.values())
    rows = [json.loads(line) for line in open(out / "metrics.jsonl")] if os.path.exists(out / "metrics.jsonl") else None
    assert rows is None or any("eval_prec" in r for r in rows)
