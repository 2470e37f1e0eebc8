"""Input pipeline: device-side ToTensor and the double-buffered
host->device prefetcher, checked against the reference's recorded pipeline outputs
(tests/golden/input_pipeline.pt) and the oracle restatement.  Bit-exact: integer -> fp32 / 255."""
import os

import pytest
import torch

from oracle import marl_oracle as O
from tests.conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_u8_to_f32_matches_reference_pipeline_vectors():
    from marlclassification_b200.input_pipeline import images_u8_to_f32

    fx = torch.load(os.path.join(GOLDEN_DIR, "input_pipeline.pt"))
    for case in fx["cases"]:
        out = images_u8_to_f32(case["u8_hwc"].to(DEV), hwc=True)
        assert torch.equal(out.cpu(), case["f32_chw"])


@pytest.mark.parametrize("b,h,w,c", [(8, 256, 256, 3), (3, 600, 600, 3), (5, 33, 17, 3), (2, 31, 64, 4), (2, 16, 24, 2),
                                     (4, 28, 28, 1), (1, 1, 1, 3), (2, 9, 12, 5)])
def test_u8_to_f32_bit_exact_vs_oracle(b, h, w, c):
    from marlclassification_b200.input_pipeline import images_u8_to_f32

    g = torch.Generator().manual_seed(b * 1000 + h + w + c)
    u8 = torch.randint(0, 256, (b, h, w, c), dtype=torch.uint8, generator=g)
    ref = O.to_tensor_batch(u8)
    out = images_u8_to_f32(u8.to(DEV), hwc=True)
    assert torch.equal(out.cpu(), ref)
    # channel-first uint8 (torchvision.io decode order): cast + divide only
    chw = u8.permute(0, 3, 1, 2).contiguous()
    assert torch.equal(images_u8_to_f32(chw.to(DEV), hwc=False).cpu(), ref)
    # caller-provided output, and an unaligned source view
    dst = torch.empty(b, c, h, w, device=DEV)
    images_u8_to_f32(u8.to(DEV), dst, hwc=True)
    assert torch.equal(dst.cpu(), ref)
    if b > 1:
        flat = torch.empty(u8.numel() + 1, dtype=torch.uint8, device=DEV)
        flat[1:] = u8.flatten().to(DEV)
        view = flat[1:].view(b, h, w, c)
        assert torch.equal(images_u8_to_f32(view, hwc=True).cpu(), ref)
        assert torch.equal(images_u8_to_f32(flat[1:].view(b, h, w, c).permute(0, 3, 1, 2).contiguous(), hwc=False).cpu(), ref)


def test_u8_to_f32_rejects_bad_inputs():
    from marlclassification_b200.input_pipeline import images_u8_to_f32

    with pytest.raises(RuntimeError):
        images_u8_to_f32(torch.zeros(1, 4, 4, 3, dtype=torch.uint8))  # CPU tensor: no CPU path
    with pytest.raises(RuntimeError):
        images_u8_to_f32(torch.zeros(1, 4, 4, 3, device=DEV))  # wrong dtype
    with pytest.raises(RuntimeError):
        images_u8_to_f32(torch.zeros(1, 4, 4, 3, dtype=torch.uint8, device=DEV), torch.empty(1, 3, 4, 5, device=DEV))


@pytest.mark.parametrize("kind", ["f32_pinned", "f32_pageable", "u8_hwc_pinned", "u8_hwc_pageable", "u8_chw", "device"])
def test_prefetcher_delivers_every_batch_in_order(kind):
    from marlclassification_b200.input_pipeline import DevicePrefetcher

    g = torch.Generator().manual_seed(3)
    n, b, c, h, w = 7, 4, 3, 20, 24
    u8 = [torch.randint(0, 256, (b, h, w, c), dtype=torch.uint8, generator=g) for _ in range(n)]
    ys = [torch.randint(0, 10, (b,), generator=g) for _ in range(n)]
    want = [O.to_tensor_batch(u) for u in u8]
    if kind.startswith("f32"):
        xs = [t.clone() for t in want]
    elif kind == "u8_chw":
        xs = [u.permute(0, 3, 1, 2).contiguous() for u in u8]
    elif kind == "device":
        xs = [t.to(DEV) for t in want]
    else:
        xs = [u.clone() for u in u8]
    if kind.endswith("pinned"):
        xs = [x.pin_memory() for x in xs]
    pf = DevicePrefetcher(list(zip(xs, ys)), DEV, hwc=(kind != "u8_chw"))
    assert len(pf) == n
    static = torch.empty(b, c, h, w, device=DEV)
    static_y = torch.empty(b, dtype=torch.int64, device=DEV)
    seen = 0
    for i, staged in enumerate(pf):
        assert staged.shape == (b, c, h, w)
        if i % 2 == 0:
            img, y = staged.deliver(static, static_y)
            assert img.data_ptr() == static.data_ptr()
        else:
            img, y = staged.deliver()
        # keep the compute stream busy so that the next copy really overlaps pending work
        busy = torch.randn(512, 512, device=DEV) @ torch.randn(512, 512, device=DEV)
        assert torch.equal(img.cpu(), want[i]) and torch.equal(y.cpu(), ys[i])
        del busy
        seen += 1
    assert seen == n


def test_prefetcher_empty_and_unlabeled():
    from marlclassification_b200.input_pipeline import DevicePrefetcher

    assert list(DevicePrefetcher([], DEV)) == []
    x = torch.rand(2, 3, 8, 8)
    (staged,) = list(DevicePrefetcher([x], DEV))
    img, y = staged.deliver()
    assert y is None and torch.equal(img.cpu(), x)
    with pytest.raises(RuntimeError):
        DevicePrefetcher([x], "cpu")


def test_train_step_from_staged_batch_equals_tensor_input():
    """The same images given as uint8 through the prefetcher or as the fp32 tensor ToTensor would
    have produced land bit-identically in the step's static input, and the trainer runs over a
    plain list 'loader' of host batches (fp32, as the reference's DataLoader yields)."""
    from marlclassification_b200.training import Trainer
    from tests.conftest import load_golden
    from tests.test_gpu_parity import run_fixture

    fx = load_golden("resisc_small")
    model, marl, env, sampler, _ = run_fixture(fx)
    nc, (_, C, H, W) = fx["model_config"]["nb_class"], fx["img"].shape
    assert C == 3
    trainer = Trainer(model, nc, 1e-3, fx["gamma"], cuda_graph=False)
    g = torch.Generator().manual_seed(5)
    u8 = [torch.randint(0, 256, (4, H, W, 3), dtype=torch.uint8, generator=g) for _ in range(3)]
    ys = [torch.randint(0, nc, (4,), generator=g) for _ in range(3)]
    for staged, u, y in zip(trainer.prefetch(list(zip(u8, ys))), u8, ys):
        out = trainer.train_step(staged, None, sampler)
        eng = sampler.engine_for(staged, gamma=fx["gamma"])
        step = trainer._Trainer__steps[id(eng)]
        assert torch.equal(step.static_img.cpu(), O.to_tensor_batch(u))
        assert torch.equal(step.static_y.cpu(), y)
        assert torch.isfinite(out[:5]).all()
    # whole-epoch API over fp32 host batches
    loader = [(O.to_tensor_batch(u), y) for u, y in zip(u8, ys)]
    before = trainer.curr_step
    trainer.train_epoch(loader, 0, sampler)
    assert trainer.curr_step == before + len(loader)
    cm = trainer.eval_epoch(loader, 0, sampler)
    assert int(cm.conf_mat().sum()) == 12
