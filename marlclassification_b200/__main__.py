"""Command line of the reference (``python -m marl_classification --run-id … train|test|infer``,
__main__.py:19-421) in front of the B200 episode: same modes, option names, defaults and
``marl.json`` / ``class_to_idx.json`` / ``models/nn_models_epoch_{e}.pt`` run-directory layout.
``--cuda`` is mandatory here (no CPU path).  Launch under ``torch.distributed.run`` to train data
parallel over the GPUs of one box."""
from __future__ import annotations

import argparse
import json
from os import makedirs
from os.path import abspath, dirname, exists, isdir, join
from typing import List, Optional, Sequence

from .registry import DATASET_REGISTRY

DEFAULT_ACTIONS = "[[1, 0], [-1, 0], [0, 1], [0, -1]]"


def parse_actions(text: str, dim: int) -> List[List[int]]:
    """``"[[1, 0], [-1, 0]]"`` -> ``[[1, 0], [-1, 0]]``; every move must be a list of ``dim``
    integers (__main__.py:329-345 validates the same with regular expressions)."""
    try:
        moves = json.loads(text)
    except json.JSONDecodeError as err:
        raise ValueError(f"Wrong action(s) : {text!r} ({err.msg})") from None
    well_formed = (isinstance(moves, list) and len(moves) > 0
                   and all(isinstance(m, list) and len(m) > 0 and all(type(v) is int for v in m) for m in moves))
    if not well_formed:
        raise ValueError(f"Wrong action(s) : {text!r} (expected a list of integer lists)")
    for i, m in enumerate(moves):
        assert len(m) == dim, f"Wrong space for action at index {i}"
    return moves


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser("Multi agent reinforcement learning for image classification - Main")
    parser.add_argument("--run-id", type=str, required=True, dest="run_id", help="run id (MLflow run name when MLflow is installed)")
    parser.add_argument("-a", "--agents", type=int, default=3, dest="agents", help="Number of agents")
    parser.add_argument("--step", type=int, default=7, help="Step number of RL episode")
    parser.add_argument("--cuda", action="store_true", dest="cuda", help="Run on CUDA (required: this build has no CPU path)")
    modes = parser.add_subparsers(dest="main_choice", required=True)

    train = modes.add_parser("train")
    train.add_argument("--action", type=str, default=DEFAULT_ACTIONS, dest="action", help="Discrete actions")
    train.add_argument("--img-size", type=int, default=28, dest="img_size", help="Image side size, assume all image are squared")
    train.add_argument("--nb-class", type=int, default=10, dest="nb_class", help="Image dataset number of class")
    train.add_argument("-d", "--dim", type=int, default=2, help="State dimension (eg. 2 -> move on a plan)")
    train.add_argument("--f", type=int, default=7, help="Window size")
    train.add_argument("--ft-extr", type=str, choices=sorted(DATASET_REGISTRY), default="mnist", dest="ft_extr_str",
                       help="Choose features extractor (CNN)")
    for flag, dest, default, text in (
        ("--nb", "n_b", 64, "Hidden size for belief LSTM"),
        ("--na", "n_a", 16, "Hidden size for action LSTM"),
        ("--nm", "n_m", 16, "Message size for NNs"),
        ("--nmo", "n_m_o", 24, "Received message output size for NNs"),
        ("--nd", "n_d", 4, "State hidden size"),
        ("--nlb", "n_l_b", 128, "Network internal hidden size for linear projections (belief unit)"),
        ("--nla", "n_l_a", 128, "Network internal hidden size for linear projections (action unit)"),
    ):
        train.add_argument(flag, type=int, default=default, dest=dest, help=text)
    train.add_argument("--res-folder", type=str, required=False,
                       default=abspath(join(dirname(abspath(__file__)), "..", "resources")),
                       help="The resources path containing the download folder with datasets")
    train.add_argument("-o", "--output-dir", type=str, required=True, dest="output_dir",
                       help="The output directory containing results and models per epoch. Created if needed.")
    train.add_argument("--batch-size", type=int, default=8, dest="batch_size",
                       help="Image batch size for training and evaluation (global batch under torchrun)")
    train.add_argument("--lr", "--learning-rate", type=float, default=1e-3, dest="learning_rate", help="learning rate")
    train.add_argument("--gamma", type=float, default=0.99, help="discount factor")
    train.add_argument("--nb-epoch", type=int, default=10, dest="nb_epoch", help="Number of training epochs")
    train.add_argument("--workers", type=int, default=6, help="DataLoader worker processes (train.py:95 uses 6)")

    test = modes.add_parser("test")
    test.add_argument("--batch-size", type=int, default=8, dest="batch_size", help="Image batch size for evaluation")
    test.add_argument("--dataset-path", type=str, required=True, dest="dataset_path", help="Input dataset path for inference")
    test.add_argument("--img-size", type=int, default=28, dest="img_size", help="Image side size, assume all image are squared")
    test.add_argument("--json-path", type=str, required=True, dest="json_path", help="JSON multi agent metadata path")
    test.add_argument("--state-dict-path", type=str, required=True, dest="state_dict_path", help="ModelsWrapper state dict path")
    test.add_argument("-o", "--output-dir", type=str, required=True, dest="output_dir",
                      help="The directory where the model outputs will be saved. Created if needed")
    test.add_argument("--workers", type=int, default=8, help="DataLoader worker processes (eval.py:57 uses 8)")

    infer = modes.add_parser("infer")
    infer.add_argument("--images", type=str, nargs="+", required=True, dest="infer_images", help="Path of images used for inference")
    infer.add_argument("--json-path", type=str, required=True, dest="json_path", help="JSON multi agent metadata path")
    infer.add_argument("--state-dict-path", type=str, required=True, dest="state_dict_path", help="ModelsWrapper state dict path")
    infer.add_argument("--class2idx", type=str, required=True, dest="class_to_idx", help="Class to index JSON file")
    infer.add_argument("-o", "--output-image-dir", type=str, required=True, dest="output_image_dir",
                       help="The directory where the model outputs will be saved. Created if needed")
    return parser


def _ensure_dir(path: str) -> None:
    if not exists(path):
        makedirs(path, exist_ok=True)
    if not isdir(path):
        raise NotADirectoryError(f'"{path}" is not a directory.')


def main(argv: Optional[Sequence[str]] = None) -> None:
    from .config import EvalConfig, InferConfig, MainConfig, ModelConfig, TrainConfig

    parser = build_parser()
    args = parser.parse_args(argv)
    main_config = MainConfig(step=args.step, run_id=args.run_id, cuda=args.cuda, nb_agent=args.agents)

    if args.main_choice == "train":
        try:
            actions = parse_actions(args.action, args.dim)
        except ValueError as err:
            parser.error(str(err))
        model_config = ModelConfig(
            ft_extr_str=args.ft_extr_str, window_size=args.f, hidden_size_belief=args.n_b, hidden_size_action=args.n_a,
            hidden_size_msg=args.n_m, hidden_size_msg_output=args.n_m_o, hidden_size_state=args.n_d, state_dim=args.dim,
            actions=actions, nb_class=args.nb_class, hidden_size_linear_belief=args.n_l_b,
            hidden_size_linear_action=args.n_l_a,
        )
        train_config = TrainConfig(
            img_size=args.img_size, nb_epoch=args.nb_epoch, learning_rate=args.learning_rate, batch_size=args.batch_size,
            resources_dir=args.res_folder, output_dir=args.output_dir, gamma=args.gamma,
        )
        _ensure_dir(args.output_dir)
        from .train import train_main

        train_main(main_config, model_config, train_config, num_workers=args.workers)
    elif args.main_choice == "test":
        eval_config = EvalConfig(
            img_size=args.img_size, state_dict_path=args.state_dict_path, batch_size=args.batch_size,
            json_path=args.json_path, dataset_path=args.dataset_path, output_dir=args.output_dir,
        )
        _ensure_dir(args.output_dir)
        from .eval import eval_main

        eval_main(main_config, eval_config, num_workers=args.workers)
    else:
        infer_config = InferConfig(
            state_dict_path=args.state_dict_path, json_path=args.json_path, images_path=args.infer_images,
            output_dir=args.output_image_dir, class_to_idx=args.class_to_idx,
        )
        _ensure_dir(args.output_image_dir)
        from .infer import infer_main

        infer_main(main_config, infer_config)


if __name__ == "__main__":
    main()
