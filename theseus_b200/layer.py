"""TheseusLayer: theseus/theseus_layer.py:29-174 (forward path).

forward(input_tensors, optimizer_kwargs) -> (Dict[name, Tensor] of optimisation variables, OptimizerInfo).
"""
from typing import Any, Dict, Optional, Tuple

import torch

from .optimizer import NonlinearLeastSquares, OptimizerInfo


class TheseusLayer(torch.nn.Module):
    def __init__(self, optimizer: NonlinearLeastSquares, vectorize: bool = True, empty_cuda_cache: bool = False):
        super().__init__()
        self.objective = optimizer.objective
        self.optimizer = optimizer

    def forward(self, input_tensors: Optional[Dict[str, torch.Tensor]] = None,
                optimizer_kwargs: Optional[Dict[str, Any]] = None) -> Tuple[Dict[str, torch.Tensor], OptimizerInfo]:
        optimizer_kwargs = optimizer_kwargs or {}
        self.objective.update(input_tensors)  # theseus_layer.py:170
        info = self.optimizer.optimize(**optimizer_kwargs)  # theseus_layer.py:171-173
        # the optimisation variables live in an engine-owned pool that the next forward() overwrites in place: hand out a snapshot
        sol = self.objective.engine().solution_tensors()
        values = dict((var.name, sol.get(var.name, var.tensor)) for var in self.objective.optim_vars.values())
        return values, info

    def to(self, *args, **kwargs):
        super().to(*args, **kwargs)
        self.objective.to(*args, **kwargs)
        return self

    @property
    def device(self):
        return self.objective.device

    @property
    def dtype(self):
        return self.objective.dtype
