"""Process setup shared by the CLI entry points: device selection (no CPU path) and, under
``torch.distributed.run``, the NCCL process group for the data-parallel step."""
from __future__ import annotations

import os

import torch as th
import torch.distributed as dist

from .parallel import DataParallelContext


def cuda_device(cuda_flag: bool) -> th.device:
    """The reference trains on CPU unless ``--cuda`` is given (train.py:77); this build has no
    CPU path, so the flag is mandatory and its absence is an error, not a silent fallback."""
    if not cuda_flag:
        raise RuntimeError("this build runs the episode on hand-written CUDA kernels only: pass --cuda "
                           "(there is no CPU path)")
    if not th.cuda.is_available():
        raise RuntimeError("--cuda given but no CUDA device is visible")
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    th.cuda.set_device(local_rank)
    return th.device("cuda", local_rank)


def data_parallel(device: th.device) -> DataParallelContext:
    """One process per GPU when launched by torchrun (WORLD_SIZE > 1); otherwise a no-op context."""
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    return DataParallelContext()


def shutdown() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
