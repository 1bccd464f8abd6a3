"""Mirror of oatomobile/torch/transforms.py (same names, same argument meaning)."""
import torch

from oatomobile_b200 import ops


def downsample_target(player_future: torch.Tensor, num_timesteps_to_keep: int) -> torch.Tensor:
  """transforms.py:23-31 — a strided view, no arithmetic."""
  _, T, _ = player_future.shape
  increments = T // num_timesteps_to_keep
  return player_future[:, 0::increments, :]


def downsample_and_transpose_visual_features(visual_features: torch.Tensor) -> torch.Tensor:
  """transforms.py:34-49 as used by dim/model.py:245-251: bilinear resize to
  100x100 (align_corners=True) fused with the H<->W transpose, one CUDA kernel."""
  return ops.transform_visual(visual_features)


def downsample_and_transpose_visual_features_hwc(visual_features: torch.Tensor) -> torch.Tensor:
  """The same for a [B,H,W,C] grid (on-disk / simulator layout, datasets/carla.py:138-140):
  the HWC->CHW permutation is folded into the kernel's loads."""
  return ops.transform_visual_hwc(visual_features)
