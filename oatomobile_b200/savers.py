"""`Checkpointer` — mirror of oatomobile/torch/savers.py:26-55 (same file naming and
format, so checkpoints are interchangeable with the reference in both directions)."""
import os

import torch


class Checkpointer:
  """Saves / restores `model.state_dict()` as `<ckpt_dir>/model-<epoch>.pt`."""

  def __init__(self, model: torch.nn.Module, ckpt_dir: str) -> None:
    os.makedirs(ckpt_dir, exist_ok=True)
    self._model = model
    self._ckpt_dir = ckpt_dir

  def save(self, epoch: int) -> str:
    """savers.py:39-46."""
    ckpt_path = os.path.join(self._ckpt_dir, "model-{}.pt".format(epoch))
    torch.save(self._model.state_dict(), ckpt_path)
    return ckpt_path

  def load(self, epoch: int) -> torch.nn.Module:
    """savers.py:48-55 — loads onto the model's current device."""
    ckpt_path = os.path.join(self._ckpt_dir, "model-{}.pt".format(epoch))
    device = next(self._model.parameters()).device
    self._model.load_state_dict(torch.load(ckpt_path, map_location=device))
    return self._model
