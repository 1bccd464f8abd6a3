"""Train-script shell with the reference's flag names and loop skeleton, on top of `Trainer`.

Mirrors oatomobile/baselines/torch/dim/train.py:38-82 (flags), :231-267 (`evaluate_epoch`),
:300-320 (epoch loop, checkpoint cadence) and the same regions of cil/train.py:

  python -m oatomobile_b200.train_script --model=dim --dataset_dir=... --output_dir=... \\
      --batch_size=512 --num_epochs=... [--save_model_frequency=4 --learning_rate=1e-3
      --num_timesteps_to_keep=4 --weight_decay=0.0 --clip_gradients]
  torchrun --nproc-per-node N -m oatomobile_b200.train_script ...    # data parallel (extension)

What differs from the reference script, and why:
* `--model dim|cil` selects what the reference keeps in two files (`dim/train.py`, `cil/train.py`).
* The step is `Trainer.train_step` (hand-written CUDA forward/backward + Adam); the reference's
  body `loss.backward(); optimizer.step()` cannot run here because the models' outputs carry no
  autograd graph.
* Batches are collated by `DeviceCollator` (pinned staging, async H2D, HWC kept) and transformed
  on the device by `model.transform` — same order as `transform()` at dim/train.py:121-135.
* The TensorBoard image logger (oatomobile/torch/loggers.py) is observability outside the hot
  path: epoch losses go to absl-style lines on stdout and `logs/losses.csv`.
* Under torchrun every rank trains on its shard of the sample list; gradients are averaged with
  one NCCL all-reduce per step; rank 0 validates, logs and checkpoints (the reference has no DDP).
"""
import argparse
import glob
import math
import os
import random
import sys
from typing import List, Sequence

import torch


def _bool(v):
  return str(v).lower() in ("1", "true", "yes", "y", "t")


def parse_flags(argv: Sequence[str]) -> argparse.Namespace:
  """The reference's absl flags (dim/train.py:38-82; names, defaults and meaning kept)."""
  p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
  p.add_argument("--dataset_dir", required=True, help="Directory of the processed dataset (with `train/` and `val/` sub-directories of .npz samples).")
  p.add_argument("--output_dir", required=True,
                 help="Where `logs/` and `ckpts/` are written.")
  p.add_argument("--batch_size", type=int, default=512,
                 help="Samples per optimiser step (split over the ranks under torchrun).")
  p.add_argument("--num_epochs", type=int, required=True,
                 help="Passes over the training set.")
  p.add_argument("--save_model_frequency", type=int, default=4,
                 help="A checkpoint is written every this many epochs.")
  p.add_argument("--learning_rate", type=float, default=1e-3, help="Adam step size.")
  p.add_argument("--num_timesteps_to_keep", type=int, default=4,
                 help="T: waypoints kept from the 80-frame future by strided down-sampling.")
  p.add_argument("--weight_decay", type=float, default=0.0,
                 help="Adam weight decay (L2).")
  p.add_argument("--clip_gradients", nargs="?", const=True, default=False, type=_bool,
                 help="Clip the global gradient norm to 1.0 before the update.")
  # extensions
  p.add_argument("--model", choices=("dim", "cil"), default="dim",
                 help="dim: ImitativeModel (dim/train.py); cil: BehaviouralModel (cil/train.py).")
  p.add_argument("--in_channels", type=int, default=2, help="BEV channels of the lidar grid.")
  p.add_argument("--seed", type=int, default=0)
  return p.parse_args(argv)


def nll_limit(T: int, noise_level: float = 1e-2) -> float:
  """Theoretical floor of the DIM loss printed next to it (dim/train.py:167-173): minus the
  log-density of N(0, noise_level * I_{2T}) at its mean."""
  d = 2 * T
  return -(-0.5 * d * math.log(2 * math.pi) - d * math.log(noise_level))


def _batches(files: List[str], batch_size: int, shuffle: bool, rng: random.Random):
  order = list(files)
  if shuffle:
    rng.shuffle(order)
  for i in range(0, len(order), batch_size):
    yield order[i:i + batch_size]


def main(argv=None) -> int:
  from oatomobile_b200 import BehaviouralModel, ImitativeModel
  from oatomobile_b200.datasets import CARLADataset, DeviceCollator
  from oatomobile_b200.savers import Checkpointer
  from oatomobile_b200.train import Trainer
  flags = parse_flags(sys.argv[1:] if argv is None else argv)
  world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local_rank)
  device = torch.device("cuda", local_rank)
  group = None
  if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=device)
    group = dist.new_group(list(range(world)))

  log_dir = os.path.join(flags.output_dir, "logs")
  ckpt_dir = os.path.join(flags.output_dir, "ckpts")
  if rank == 0:
    os.makedirs(log_dir, exist_ok=True)
    os.makedirs(ckpt_dir, exist_ok=True)

  torch.manual_seed(flags.seed)  # every replica builds the same initial weights (+ Trainer syncs)
  output_shape = (flags.num_timesteps_to_keep, 2)
  cls = ImitativeModel if flags.model == "dim" else BehaviouralModel
  model = cls(output_shape=output_shape, in_channels=flags.in_channels).to(device)
  trainer = Trainer(model, lr=flags.learning_rate, weight_decay=flags.weight_decay,
                    clip_gradients=flags.clip_gradients, noise_level=1e-2, group=group)
  checkpointer = Checkpointer(model=model, ckpt_dir=ckpt_dir)

  modalities = ("lidar", "is_at_traffic_light", "traffic_light_state", "player_future", "velocity")
  with_mode = flags.model == "cil"  # cil/train.py:137-149 asks the loader for the command label
  all_train = sorted(glob.glob(os.path.join(flags.dataset_dir, "train", "*.npz")))
  # equal shards: every rank must take the same number of optimiser steps (one all-reduce each)
  train_files = all_train[rank::world][:len(all_train) // world]
  val_files = sorted(glob.glob(os.path.join(flags.dataset_dir, "val", "*.npz")))
  if not train_files:
    raise SystemExit("no training samples under %s" % os.path.join(flags.dataset_dir, "train"))
  per_rank = max(flags.batch_size // world, 1)
  collate = DeviceCollator(device)
  rng = random.Random(flags.seed + rank)

  def load(files):
    samples = [CARLADataset.load_datum(f, modalities, mode=with_mode, dataformat="HWC") for f in files]
    return model.transform(collate(samples))  # dim/train.py:121-135

  floor = nll_limit(output_shape[0]) if flags.model == "dim" else None
  csv = open(os.path.join(log_dir, "losses.csv"), "a") if rank == 0 else None
  for epoch in range(flags.num_epochs):  # dim/train.py:300-320
    total, n = torch.zeros((), device=device), 0
    for files in _batches(train_files, per_rank, True, rng):
      total = total + trainer.train_step(load(files)) * len(files)
      n += len(files)
    loss_train = float(total) / max(n, 1)
    loss_val = float("nan")
    if rank == 0 and val_files:  # dim/train.py:231-267: eval-mode loss, no parameter update
      vt, vn = 0.0, 0
      for files in _batches(val_files, per_rank * 5, False, rng):
        vt += float(trainer.evaluate_step(load(files))) * len(files)
        vn += len(files)
      loss_val = vt / max(vn, 1)
    if rank == 0:
      extra = "" if floor is None else " | THEORETICAL MIN: %.2f" % floor
      print("EPOCH %d | TRAIN LOSS: %.4f | VAL LOSS: %.4f%s" % (epoch, loss_train, loss_val, extra), flush=True)
      csv.write("%d,%.8f,%.8f\n" % (epoch, loss_train, loss_val))
      csv.flush()
      if epoch % flags.save_model_frequency == 0:  # dim/train.py:310-312
        checkpointer.save(epoch)
  if world > 1:
    import torch.distributed as dist
    dist.destroy_process_group()
  return 0


if __name__ == "__main__":
  sys.exit(main())
