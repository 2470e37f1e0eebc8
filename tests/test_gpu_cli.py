"""The reference's CLI workflow end to end on the GPU (README usage: train -> test -> infer):
``python -m marl_classification --run-id … --cuda train|test|infer`` on a synthetic MNIST-shaped
image folder, checking the run-directory layout the reference produces (train.py:30-147,
eval.py, infer.py) and that the saved ``state_dict`` loads back under the reference's key names."""
import json
import os

import pytest
import torch
from PIL import Image

from tests.clidata import make_image_folder
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu


def test_train_test_infer_roundtrip(tmp_path, capsys):
    from marlclassification_b200.__main__ import main

    res = tmp_path / "resources"
    root = make_image_folder(str(res / "downloaded" / "mnist_png" / "all_png"), classes=3, per_class=17, size=28,
                             grey_every=1)  # 51 grey PNGs -> 43 train (5 x 8 + ragged 3) / 8 eval
    out = tmp_path / "run"
    common = ["--run-id", "cli_test", "-a", "3", "--step", "4", "--cuda"]
    main(common + ["train", "--ft-extr", "mnist", "--f", "6", "--nb", "64", "--na", "64", "--nm", "16", "--nmo", "24",
                   "--nd", "8", "--nlb", "96", "--nla", "96", "--nb-class", "3", "--batch-size", "8", "--nb-epoch", "2",
                   "--res-folder", str(res), "-o", str(out), "--workers", "0"])
    files = set(os.listdir(out))
    assert {"marl.json", "class_to_idx.json", "models", "confusion_matrix_epoch_0_eval.png",
            "confusion_matrix_epoch_1_eval.png", "pred_original.png", "pred_step_3.png", "animated_gif.gif"} <= files
    assert sorted(os.listdir(out / "models")) == ["nn_models_epoch_0.pt", "nn_models_epoch_1.pt"]
    assert json.load(open(out / "class_to_idx.json")) == {"class_0": 0, "class_1": 1, "class_2": 2}
    marl = json.load(open(out / "marl.json"))
    assert marl["window_size"] == 6 and marl["actions"] == [[1, 0], [-1, 0], [0, 1], [0, -1]] and marl["nb_class"] == 3
    if "metrics.jsonl" in files:  # JSONL tracker (MLflow absent): step-0 train metrics + one eval row per epoch
        rows = [json.loads(line) for line in open(out / "metrics.jsonl")]
        assert any("loss" in r for r in rows) and sum("eval_prec" in r for r in rows) == 2
    sd = torch.load(out / "models" / "nn_models_epoch_1.pt", map_location="cpu")
    ref_keys = list(load_golden("mnist_ckpt")["state_dict"])
    assert list(sd) == ref_keys  # the reference's (name-mangled) keys, in its order
    sd0 = torch.load(out / "models" / "nn_models_epoch_0.pt", map_location="cpu")
    assert any(not torch.equal(sd[k], sd0[k]) for k in sd) and all(torch.isfinite(v).all() for v in sd.values())

    capsys.readouterr()
    main(common + ["test", "--dataset-path", root, "--json-path", str(out / "marl.json"), "--state-dict-path",
                   str(out / "models" / "nn_models_epoch_1.pt"), "--batch-size", "8", "-o", str(tmp_path / "test_out"),
                   "--workers", "0"])
    printed = capsys.readouterr().out
    assert "Precision mean" in printed and "Recall mean" in printed and '"class_2"' in printed
    assert os.path.exists(tmp_path / "test_out" / "confusion_matrix_epoch_0_test.png")

    img_path = os.path.join(root, "class_1", "img_0.png")
    main(common + ["infer", "--images", img_path, "--json-path", str(out / "marl.json"), "--state-dict-path",
                   str(out / "models" / "nn_models_epoch_1.pt"), "--class2idx", str(out / "class_to_idx.json"),
                   "-o", str(tmp_path / "infer_out")])
    inf = tmp_path / "infer_out" / "img_0.png"
    assert open(inf / "info.txt").read().splitlines()[0] == img_path
    assert Image.open(inf / "animated_gif.gif").n_frames >= 2


def test_cli_eval_matches_episode_api(tmp_path):
    """`test` mode = EpisodeSampler.run_episode_get_last_step + mean over agents + ConfusionMeter
    (eval.py:66-75): with the shipped MNIST checkpoint both give the same number of samples and a
    confusion matrix whose entries sum to the dataset size."""
    from marlclassification_b200.config import EvalConfig, MainConfig, ModelConfig
    from marlclassification_b200.eval import eval_main

    fx = load_golden("mnist_ckpt")
    root = make_image_folder(str(tmp_path / "imgs"), classes=10, per_class=2, size=28, grey_every=1)
    cfg = ModelConfig(**fx["model_config"])
    cfg.save_marl_config(str(tmp_path / "marl.json"))
    torch.save(fx["state_dict"], tmp_path / "sd.pt")
    meter = eval_main(MainConfig(step=5, run_id="x", cuda=True, nb_agent=3),
                      EvalConfig(img_size=28, state_dict_path=str(tmp_path / "sd.pt"), batch_size=8,
                                 json_path=str(tmp_path / "marl.json"), dataset_path=root, output_dir=str(tmp_path / "o")),
                      num_workers=0)
    cm = meter.conf_mat()
    assert cm.shape == (10, 10) and int(cm.sum()) == 20
    assert int(cm.sum(dim=1).max()) == 2  # two images per true class
