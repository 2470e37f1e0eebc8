"""Shared pytest config: the ``gpu`` marker, golden-fixture loading, helpers."""
import glob
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_NAMES = sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.pt")))
                      if n not in ("input_pipeline", "metrics"))  # episode fixtures; the others hold ToTensor / meter vectors


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), map_location="cpu", weights_only=False)


def oracle_config(mc):
    """tests/golden model_config (marl.json schema, config.py:18-31) -> OracleConfig."""
    from oracle.marl_oracle import OracleConfig

    return OracleConfig(
        ft_extr=mc["ft_extr_str"], f=mc["window_size"], n_b=mc["hidden_size_belief"],
        n_a=mc["hidden_size_action"], n_m=mc["hidden_size_msg"], n_m_o=mc["hidden_size_msg_output"],
        n_d=mc["hidden_size_state"], nb_class=mc["nb_class"], nl_b=mc["hidden_size_linear_belief"],
        nl_a=mc["hidden_size_linear_action"], actions=mc["actions"],
    )


def rel_l2(a, b):
    """Norm-wise relative error ||a-b|| / ||b|| (element-wise
    ratios are meaningless for logits / values that cross zero)."""
    a, b = a.double().flatten(), b.double().flatten()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


@pytest.fixture(params=GOLDEN_NAMES)
def golden(request):
    return load_golden(request.param)
