"""CPU checks of the CLI layer (reference: __main__.py, train.py, eval.py, infer.py,
data/datasets.py, visualization.py): option surface, image-folder dataset against torchvision's
ImageFolder + ToTensor (what the reference uses), data-parallel batch sampler, tracker and the
PIL-drawn pictures.  Nothing here runs the episode (that is tests/test_gpu_cli.py)."""
import json
import os

import numpy as np
import pytest
import torch
from PIL import Image

from marlclassification_b200.__main__ import DEFAULT_ACTIONS, build_parser, main, parse_actions
from marlclassification_b200.data import (
    FolderDataset, ShardedBatchSampler, collate_images, default_image_pipeline, u8_image_pipeline,
)
from marlclassification_b200.registry import DATASET_REGISTRY, get_dataset_spec
from tests.clidata import make_image_folder


def test_parser_defaults_match_reference_cli():
    """Defaults of __main__.py:47-219 (agents 3, step 7, f 7, nb 64, na 16, nm 16, nmo 24, nd 4,
    nlb = nla = 128, batch 8, lr 1e-3, gamma 0.99, 10 epochs, mnist)."""
    a = build_parser().parse_args(["--run-id", "r", "train", "-o", "out"])
    assert (a.agents, a.step, a.cuda, a.main_choice) == (3, 7, False, "train")
    assert (a.f, a.dim, a.n_b, a.n_a, a.n_m, a.n_m_o, a.n_d, a.n_l_b, a.n_l_a) == (7, 2, 64, 16, 16, 24, 4, 128, 128)
    assert (a.batch_size, a.learning_rate, a.gamma, a.nb_epoch, a.ft_extr_str, a.nb_class, a.img_size) == (
        8, 1e-3, 0.99, 10, "mnist", 10, 28)
    assert parse_actions(a.action, a.dim) == [[1, 0], [-1, 0], [0, 1], [0, -1]] and a.action == DEFAULT_ACTIONS
    t = build_parser().parse_args(["--run-id", "r", "-a", "5", "--step", "9", "--cuda", "test", "--dataset-path", "d",
                                   "--json-path", "j", "--state-dict-path", "s", "-o", "o"])
    assert (t.agents, t.step, t.cuda, t.batch_size, t.main_choice) == (5, 9, True, 8, "test")
    i = build_parser().parse_args(["--run-id", "r", "infer", "--images", "a.png", "b/*.png", "--json-path", "j",
                                   "--state-dict-path", "s", "--class2idx", "c", "-o", "o"])
    assert i.infer_images == ["a.png", "b/*.png"] and i.output_image_dir == "o"


def test_parser_requires_run_id_mode_and_output():
    for argv in (["train", "-o", "x"], ["--run-id", "r"], ["--run-id", "r", "train"],
                 ["--run-id", "r", "train", "-o", "x", "--ft-extr", "imagenet"]):
        with pytest.raises(SystemExit):
            build_parser().parse_args(argv)


@pytest.mark.parametrize("text", ["[[1,0],[1]]x", "[1,0]", "[[1.5,0]]", "[]", "[[true,0]]", "[[]]"])
def test_parse_actions_rejects_malformed(text):
    with pytest.raises(ValueError, match="Wrong action"):
        parse_actions(text, 2)


def test_parse_actions_checks_dimension():
    assert parse_actions("[[3,0],[-3,0],[0,0]]", 2) == [[3, 0], [-3, 0], [0, 0]]
    with pytest.raises(AssertionError, match="index 1"):
        parse_actions("[[1,0],[1,0,0]]", 2)


def test_cli_without_cuda_flag_fails_loudly(tmp_path):
    """The reference falls back to CPU without --cuda; this build must refuse (no CPU path)."""
    with pytest.raises(RuntimeError, match="no CPU path"):
        main(["--run-id", "r", "train", "-o", str(tmp_path / "out"), "--res-folder", str(tmp_path)])


def test_registry_has_reference_names_and_both_constructors(tmp_path):
    assert sorted(DATASET_REGISTRY) == ["aid", "kneemri", "mnist", "resisc45", "skin_cancer", "worldstrat"]
    with pytest.raises(AssertionError, match="Unknown dataset"):
        get_dataset_spec("cifar")
    with pytest.raises(NotImplementedError):
        get_dataset_spec("kneemri").dataset_constructor(str(tmp_path), default_image_pipeline())
    with pytest.raises(AssertionError, match="does not exist"):
        get_dataset_spec("aid").dataset_constructor(str(tmp_path), default_image_pipeline())
    root = make_image_folder(str(tmp_path / "downloaded" / "mnist_png" / "all_png"), classes=3, per_class=2, size=28)
    ds = get_dataset_spec("mnist").dataset_constructor(str(tmp_path), default_image_pipeline())
    assert ds.root == root and len(ds) == 6


def test_folder_dataset_equals_torchvision_imagefolder(tmp_path):
    """Same class_to_idx, same sample order, bit-identical pixels as ImageFolder + ToTensor with the
    reference's RGB loader (datasets.py:16-43, registry.py:56-57), grey-scale PNGs included."""
    tv = pytest.importorskip("torchvision")
    root = make_image_folder(str(tmp_path / "imgs"), classes=4, per_class=3, size=20, grey_every=2, nested=True)
    ours = FolderDataset(root, default_image_pipeline())
    ref = tv.datasets.ImageFolder(root, transform=tv.transforms.Compose([tv.transforms.ToTensor()]),
                                  loader=lambda p: Image.open(p).convert("RGB"))
    assert ours.class_to_idx == ref.class_to_idx
    assert ours.samples == ref.samples and ours.targets == ref.targets
    raw = FolderDataset(root, u8_image_pipeline())
    for k in range(len(ref)):
        (xo, yo), (xr, yr), (xb, _) = ours[k], ref[k], raw[k]
        assert yo == yr and xo.dtype == torch.float32 and torch.equal(xo, xr)
        assert xb.dtype == torch.uint8 and xb.shape == (20, 20, 3)
        assert torch.equal(xb.permute(2, 0, 1).float().div(255), xr)  # the conversion the device kernel performs


def test_folder_dataset_errors(tmp_path):
    with pytest.raises(AssertionError):
        FolderDataset(str(tmp_path / "missing"), default_image_pipeline())
    with pytest.raises(FileNotFoundError):
        FolderDataset(str(tmp_path), default_image_pipeline())
    os.makedirs(tmp_path / "c0")
    (tmp_path / "c0" / "notes.txt").write_text("x")
    with pytest.raises(FileNotFoundError):
        FolderDataset(str(tmp_path), default_image_pipeline())


def test_collate_images_stacks_bytes_and_labels():
    items = [(torch.full((5, 4, 3), k, dtype=torch.uint8), k) for k in range(3)]
    x, y = collate_images(items)
    assert x.shape == (3, 5, 4, 3) and x.dtype == torch.uint8 and y.dtype == torch.int64 and y.tolist() == [0, 1, 2]


@pytest.mark.parametrize("length,batch,world", [(37, 8, 1), (37, 8, 2), (64, 16, 4), (5, 8, 2), (3, 4, 4), (0, 4, 2)])
def test_sharded_batch_sampler_partitions_each_global_batch(length, batch, world):
    """Ranks agree on the global batches, take disjoint equal slices of each, and together cover
    every index except the (< world) images trimmed from a ragged last batch."""
    samplers = [ShardedBatchSampler(length, batch, r, world, shuffle=True, seed=3) for r in range(world)]
    per_rank = [list(s) for s in samplers]
    assert len({len(b) for b in per_rank}) == 1 and all(len(b) == len(samplers[0]) for b in per_rank)
    seen = []
    for step in zip(*per_rank):
        assert len({len(s) for s in step}) == 1  # equal shards -> grad average == global mean
        merged = [i for s in step for i in s]
        assert len(merged) <= batch and len(set(merged)) == len(merged)
        seen += merged
    assert len(set(seen)) == len(seen) and set(seen) <= set(range(length))
    full, rest = divmod(length, batch)
    assert len(seen) == full * batch + rest - rest % world
    if world == 1:
        assert sorted(seen) == list(range(length))  # drop_last=False, like the reference's loaders


def test_sharded_batch_sampler_epochs_and_validation():
    s = ShardedBatchSampler(50, 10, 0, 1, shuffle=True, seed=1)
    e0 = list(s)
    assert list(s) == e0  # deterministic within an epoch: every rank can rebuild it
    s.set_epoch(1)
    assert list(s) != e0 and sorted(i for b in s for i in b) == list(range(50))
    assert list(ShardedBatchSampler(5, 2, shuffle=False)) == [[0, 1], [2, 3], [4]]
    with pytest.raises(ValueError):
        ShardedBatchSampler(10, 6, 0, 4)
    with pytest.raises(ValueError):
        ShardedBatchSampler(10, 8, 4, 4)


def test_run_tracker_jsonl_backend(tmp_path):
    from marlclassification_b200.tracking import RunTracker

    tr = RunTracker("MARLClassification", "train_x", str(tmp_path), use_mlflow=False)
    assert tr.backend == "jsonl"
    tr.log_param("device", "cuda")
    tr.log_params({"step": 5, "actions": [[1, 0]]})
    tr.log_metrics(0, {"loss": 1.5})
    tr.log_metrics(100, {"loss": 0.5, "train_prec": 0.25})
    tr.end()
    params = json.load(open(tmp_path / "params.json"))
    assert params["run_name"] == "train_x" and params["device"] == "cuda" and params["actions"] == [[1, 0]]
    rows = [json.loads(line) for line in open(tmp_path / "metrics.jsonl")]
    assert [r["step"] for r in rows] == [0, 100] and rows[1]["train_prec"] == 0.25


def test_confusion_matrix_png_and_visualised_episode(tmp_path):
    """File names of metrics.py:110-129 and visualization.py:45-99; the episode is a stub (the
    real one needs the GPU), which pins what visualize_steps consumes: step_preds / step_pos."""
    from marlclassification_b200.core.episode import EpisodeDetailedOutput
    from marlclassification_b200.metrics import ConfusionMeter
    from marlclassification_b200.visualization import visualize_steps

    cm = ConfusionMeter(3)
    cm.add(torch.eye(3)[[0, 1, 1, 2]], torch.tensor([0, 1, 2, 2]))
    cm.save_conf_matrix(4, str(tmp_path), "eval")
    assert Image.open(tmp_path / "confusion_matrix_epoch_4_eval.png").size[0] > 0

    T, Na, f, H, W = 3, 2, 4, 12, 10

    class StubSampler:
        def run_episode(self, img):
            assert img.shape == (1, 1, H, W)
            preds = torch.zeros(T, Na, 1, 3)
            preds[:, :, :, 2] = 5.0
            pos = torch.tensor([[[[t, t + a]] for a in range(Na)] for t in range(T)])
            return EpisodeDetailedOutput(preds, torch.zeros(T, Na, 1), torch.zeros(T, Na, 1), pos)

    img = torch.rand(1, H, W)
    visualize_steps(StubSampler(), img, img, f, str(tmp_path), {"zero": 0, "one": 1, "two": 2})
    names = sorted(os.listdir(tmp_path))
    assert {"pred_original.png", "pred_step_0.png", "pred_step_2.png", "animated_gif.gif"} <= set(names)
    gif = Image.open(tmp_path / "animated_gif.gif")
    # PIL folds the five identical "original" frames into one shown 5 x 200 ms
    assert gif.n_frames == 1 + T and gif.info["duration"] == 1000
    # the last frame shows exactly the uncovered windows: pixel (0,0) was seen at t=0 by agent 0
    frame = np.asarray(Image.open(tmp_path / "pred_step_2.png").convert("RGB"))
    assert frame.shape[0] > H and frame.shape[1] >= W


def test_split_is_seeded_disjoint_and_85_15():
    from marlclassification_b200.train import split_indices

    tr, te = split_indices(51)
    assert (len(tr), len(te)) == (43, 8) and sorted(tr + te) == list(range(51))
    assert split_indices(51) == (tr, te)  # every data-parallel rank cuts the same way
    assert split_indices(0) == ([], [])


def test_load_trained_names_the_missing_file(tmp_path):
    from marlclassification_b200.config import ModelConfig

    with pytest.raises(AssertionError, match="JSON path"):
        ModelConfig.load_trained(str(tmp_path / "marl.json"), str(tmp_path / "sd.pt"), 3, "cuda")
    (tmp_path / "marl.json").write_text("{}")
    with pytest.raises(AssertionError, match="State dict path"):
        ModelConfig.load_trained(str(tmp_path / "marl.json"), str(tmp_path / "sd.pt"), 3, "cuda")


def test_multi_agent_step_protocol_with_injected_samples(monkeypatch):
    """MultiAgent.act (agent.py:40-68) against a stub network on the CPU: the state and the message
    are threaded from step to step, the draw goes through ``torch.multinomial`` with the
    reference's call form (so patching it injects samples), log-probs are log(probs[action])."""
    from marlclassification_b200.core.agent import AgentOutput, MultiAgent
    from marlclassification_b200.networks.models import ModelOutput, RecurrentOutput

    na, nb, n_act = 2, 3, 4
    calls = []

    class StubNet:
        nb_class, device = 5, torch.device("cpu")

        def random_first_state(self, a, b):
            return RecurrentOutput(*(torch.full((a, b, 2), float(k)) for k in range(4)))

        def zero_first_message(self, a, b):
            return torch.zeros(a, b, 1)

        def __call__(self, obs, msg, npos, hidden):
            calls.append((msg.clone(), hidden.h.clone()))
            probs = torch.softmax(torch.arange(na * nb * n_act, dtype=torch.float32).view(na, nb, n_act) / 7, -1)
            out = ModelOutput(actions_probabilities=probs, values=torch.ones(na, nb),
                              predictions=torch.zeros(na, nb, 5), messages=msg + 1)
            return out, RecurrentOutput(hidden.h + 10, hidden.c, hidden.h_caret, hidden.c_caret)

    agents = MultiAgent(na, StubNet())
    assert len(agents) == na and agents.nb_class == 5 and agents.device.type == "cpu"
    with pytest.raises(AssertionError, match="reset"):
        agents.act(torch.zeros(na, nb, 1, 2, 2), torch.zeros(na, nb, 2))
    agents.reset(nb)

    forced = torch.tensor([[3], [0], [1], [2], [2], [0]])
    seen = {}

    def fake_multinomial(p, num_samples, replacement):
        seen["shape"], seen["kw"] = tuple(p.shape), (num_samples, replacement)
        return forced

    monkeypatch.setattr(torch, "multinomial", fake_multinomial)
    out = agents.act(torch.zeros(na, nb, 1, 2, 2), torch.zeros(na, nb, 2))
    assert isinstance(out, AgentOutput) and seen == {"shape": (na * nb, n_act), "kw": (1, True)}
    assert torch.equal(out.actions, forced.view(na, nb))
    probs = torch.softmax(torch.arange(na * nb * n_act, dtype=torch.float32).view(na, nb, n_act) / 7, -1)
    assert torch.equal(out.actions_log_probs, probs.gather(-1, forced.view(na, nb, 1)).squeeze(-1).log())
    agents.act(torch.zeros(na, nb, 1, 2, 2), torch.zeros(na, nb, 2))
    assert calls[0][0].eq(0).all() and calls[1][0].eq(1).all()      # the message of step 0 is heard at step 1
    assert calls[0][1].eq(0).all() and calls[1][1].eq(10).all()     # and so is the recurrent state


@pytest.mark.parametrize("nb_class", [2, 7, 31])
def test_reference_metrics_scenarios(tmp_path, nb_class):
    """The two scenarios of the reference's tests/test_metrics.py, through the alias package: identity
    predictions with one planted error, picture written; mean of a loss window."""
    from marl_classification.metrics import ConfusionMeter, LossMeter

    y_pred = torch.eye(nb_class)
    y_pred[0, 0], y_pred[0, 1] = 0.0, 1.0  # sample 0 (class 0) is predicted as class 1
    meter = ConfusionMeter(nb_class, None)
    meter.add(y_pred, torch.arange(nb_class))
    cm = meter.conf_mat()
    assert cm[0, 0] == 0 and cm[0, 1] == 1 and (torch.diag(cm)[1:] == 1).all() and cm.sum() == nb_class
    assert meter.recall()[0] == 0 and meter.precision()[1] == 0.5
    meter.save_conf_matrix(0, str(tmp_path), "unittest")
    assert (tmp_path / "confusion_matrix_epoch_0_unittest.png").stat().st_size > 0
    losses = LossMeter(None)
    for v in (0.5, 0.25, 0.75, 0.5):
        losses.add(v)
    assert losses.loss() == 0.5


@pytest.mark.parametrize("window", [None, 1, 4])
def test_running_confusion_matrix_equals_rebuild_from_window(window):
    """The meter keeps a running count (add the new batch, subtract the one leaving the window);
    the reference rebuilds from the stored window on every query (metrics.py:56-108).  Same matrix,
    precision, recall and packed means after every batch, ragged batch sizes included."""
    from marlclassification_b200.metrics import ConfusionMeter

    nc, g = 5, torch.Generator().manual_seed(7)
    meter, kept = ConfusionMeter(nc, window), []
    assert meter.conf_mat().sum() == 0
    for step in range(12):
        b = int(torch.randint(1, 9, (1,), generator=g))
        proba, true = torch.rand(b, nc, generator=g), torch.randint(nc, (b,), generator=g)
        meter.add(proba, true)
        kept.append((proba.argmax(1), true))
        if window is not None:
            kept = kept[-window:]
        pred, tgt = torch.cat([p for p, _ in kept]), torch.cat([t for _, t in kept])
        want = torch.zeros(nc, nc, dtype=torch.int64)
        for p, t in zip(pred.tolist(), tgt.tolist()):
            want[t, p] += 1
        assert torch.equal(meter.conf_mat(), want)
        wf = want.float()
        prec = torch.where(wf.sum(0) != 0, wf.diag() / wf.sum(0).clamp(min=1), torch.zeros(nc))
        rec = torch.where(wf.sum(1) != 0, wf.diag() / wf.sum(1).clamp(min=1), torch.zeros(nc))
        assert torch.equal(meter.precision(), prec) and torch.equal(meter.recall(), rec)
        assert torch.allclose(meter.mean_precision_recall(), torch.stack((prec.mean(), rec.mean())))


def test_resident_split_serves_the_same_batches_as_the_streaming_loader(tmp_path):
    """data.ResidentImages (decode once, index in device memory) against a DataLoader over the same
    sampler: identical bytes and labels batch for batch, across an epoch change, ragged tail
    included.  (Device-agnostic torch indexing, exercised here on the CPU.)"""
    from torch.utils.data import DataLoader, Subset

    from marlclassification_b200.data import ResidentImages

    root = make_image_folder(str(tmp_path / "imgs"), classes=3, per_class=7, size=12, grey_every=3)
    ds = FolderDataset(root, u8_image_pipeline())
    indices = [20, 3, 5, 8, 13, 1, 0, 19, 7, 11, 2]
    for world, rank in ((1, 0), (2, 1)):
        sampler = ShardedBatchSampler(len(indices), 4, rank, world, shuffle=True, seed=5)
        resident = ResidentImages(ds, indices, sampler, "cpu", decode_threads=3)
        assert resident.images.shape == (11, 12, 12, 3) and resident.nbytes == 11 * 12 * 12 * 3
        stream = DataLoader(Subset(ds, indices), batch_sampler=sampler, collate_fn=collate_images)
        for epoch in (0, 1):
            sampler.set_epoch(epoch)
            got, want = list(resident), list(stream)
            assert len(got) == len(want) == len(resident) > 0
            for (xg, yg), (xw, yw) in zip(got, want):
                assert xg.dtype == torch.uint8 and torch.equal(xg, xw) and torch.equal(yg, yw)


def test_resident_split_rejects_mixed_sizes_and_foreign_samplers(tmp_path):
    from marlclassification_b200.data import ResidentImages

    root = make_image_folder(str(tmp_path / "imgs"), classes=2, per_class=2, size=8)
    Image.fromarray(np.zeros((9, 8, 3), np.uint8), "RGB").save(os.path.join(root, "class_1", "odd.png"))
    ds = FolderDataset(root, u8_image_pipeline())
    with pytest.raises(ValueError, match="equally sized"):
        ResidentImages(ds, list(range(len(ds))), ShardedBatchSampler(len(ds), 2), "cpu")
    with pytest.raises(ValueError, match="batch sampler covers"):
        ResidentImages(ds, [0, 1], ShardedBatchSampler(3, 2), "cpu")
    empty = ResidentImages(ds, [], ShardedBatchSampler(0, 2), "cpu")
    assert len(empty) == 0 and list(empty) == []
