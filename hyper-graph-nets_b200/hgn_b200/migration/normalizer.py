"""Mirror of src/migration/normalizer.py: online mean/std accumulator (plain attributes, like the
reference: the statistics are NOT in ``state_dict``, they travel with the pickled object)."""
import torch
from torch import nn, Tensor

from .. import util


class Normalizer(nn.Module):
    """Feature normalizer that accumulates statistics online."""

    def __init__(self, size: int, name: str, max_accumulations=10 ** 6, std_epsilon=1e-8) -> None:
        super().__init__()
        dev = util.device
        self._name = name
        self._max_accumulations = max_accumulations
        self._std_epsilon = torch.tensor([std_epsilon], device=dev)
        self._acc_count = torch.zeros(1, dtype=torch.float32, device=dev)
        self._num_accumulations = torch.zeros(1, dtype=torch.float32, device=dev)
        self._acc_sum = torch.zeros(size, dtype=torch.float32, device=dev)
        self._acc_sum_squared = torch.zeros(size, dtype=torch.float32, device=dev)

    def forward(self, batched_data: Tensor, accumulate=True) -> Tensor:
        """Normalizes input data and accumulates statistics (at most ``max_accumulations`` times).

        While a CUDA graph is being captured (``hgn_b200.graphed``) the accumulation becomes part of the graph: it updates the
        statistics IN PLACE -- every replay must add to what the previous replay left -- and the reference's host-side comparison
        of the accumulation counter (a device read-back, normalizer.py:44) is not made: a captured step always accumulates (the
        limit is 10^6 accumulations, about 50 000 trajectories)."""
        if accumulate:
            if batched_data.is_cuda and torch.cuda.is_current_stream_capturing():
                self._accumulate_in_place(batched_data)
            elif self._num_accumulations < self._max_accumulations:
                self._accumulate(batched_data)
        return (batched_data - self._mean()) / self._std_with_epsilon()

    def inverse(self, normalized_batch_data: Tensor) -> Tensor:
        return normalized_batch_data * self._std_with_epsilon() + self._mean()

    def _accumulate(self, batched_data: Tensor) -> None:
        rows = torch.tensor(batched_data.shape[0], dtype=torch.float32, device=self._acc_count.device)
        self._acc_sum = self._acc_sum.add(torch.sum(batched_data, dim=0))
        self._acc_sum_squared = self._acc_sum_squared.add(torch.sum(batched_data ** 2, dim=0))
        self._acc_count = self._acc_count.add(rows)
        self._num_accumulations = self._num_accumulations.add(1.)

    def _accumulate_in_place(self, batched_data: Tensor) -> None:
        self._acc_sum.add_(torch.sum(batched_data, dim=0))
        self._acc_sum_squared.add_(torch.sum(batched_data ** 2, dim=0))
        self._acc_count.add_(float(batched_data.shape[0]))
        self._num_accumulations.add_(1.)

    def _safe_count(self) -> Tensor:
        return torch.maximum(self._acc_count, torch.ones_like(self._acc_count))

    def _mean(self) -> Tensor:
        return self._acc_sum / self._safe_count()

    def _std_with_epsilon(self) -> Tensor:
        var = torch.abs(self._acc_sum_squared / self._safe_count() - self._mean() ** 2)
        return torch.maximum(torch.sqrt(var), self._std_epsilon)

    def get_acc_sum(self) -> Tensor:
        return self._acc_sum
