"""Experiment tracking for the CLI.  The reference logs to MLflow (train.py:41-49, 111-143);
MLflow is not part of this image, so the tracker uses it when it imports and otherwise writes the
same parameters / metrics as JSON lines under the run's output directory."""
from __future__ import annotations

import json
import time
from os.path import join
from typing import Any, Dict, Optional


class RunTracker:
    def __init__(self, experiment: str, run_name: str, output_dir: str, use_mlflow: Optional[bool] = None) -> None:
        self._mlflow = None
        if use_mlflow is None or use_mlflow:
            try:
                import mlflow  # type: ignore

                self._mlflow = mlflow
            except ImportError:
                if use_mlflow:
                    raise
        self._params_path = join(output_dir, "params.json")
        self._metrics_path = join(output_dir, "metrics.jsonl")
        self._params: Dict[str, Any] = {"experiment": experiment, "run_name": run_name}
        if self._mlflow is not None:
            self._mlflow.set_experiment(experiment)
            self._mlflow.start_run(run_name=run_name)
        else:
            open(self._metrics_path, "w", encoding="utf-8").close()
            self._flush_params()

    @property
    def backend(self) -> str:
        return "mlflow" if self._mlflow is not None else "jsonl"

    def _flush_params(self) -> None:
        with open(self._params_path, "w", encoding="utf-8") as fh:
            json.dump(self._params, fh, default=str)

    def log_param(self, key: str, value: Any) -> None:
        self.log_params({key: value})

    def log_params(self, params: Dict[str, Any]) -> None:
        if self._mlflow is not None:
            self._mlflow.log_params(params)
        else:
            self._params.update(params)
            self._flush_params()

    def log_metrics(self, step: int, metrics: Dict[str, float]) -> None:
        if self._mlflow is not None:
            self._mlflow.log_metrics(step=step, metrics=metrics)
        else:
            with open(self._metrics_path, "a", encoding="utf-8") as fh:
                fh.write(json.dumps({"step": step, "time": time.time(), **metrics}) + "\n")

    def end(self) -> None:
        if self._mlflow is not None:
            self._mlflow.end_run()
