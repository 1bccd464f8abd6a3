"""TEST TOOLING — builds and drives tests/emu/train_emu.cpp: the training-step work-item
bodies of oatomobile_b200/csrc/train_functors.h executed on the host by a plain loop.
Used by the CPU unit tests (math of every kernel body against the oracle without a GPU)
and by the GPU tests as a second, bit-near reference for the CUDA launch plumbing."""
import ctypes
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libtrain_emu.so")
DEPS = [os.path.join(HERE, "train_emu.cpp"),
        os.path.join(ROOT, "oatomobile_b200", "csrc", "train_functors.h"),
        os.path.join(ROOT, "oatomobile_b200", "csrc", "train_impl.h"),
        os.path.join(ROOT, "include", "oat_b200.h")]


class TT(ctypes.Structure):
  _fields_ = [("name", ctypes.c_char_p), ("param", ctypes.c_void_p), ("grad", ctypes.c_void_p),
              ("ndim", ctypes.c_int32), ("shape", ctypes.c_int64 * 4)]


_lib = None


def lib():
  global _lib
  if _lib is None:
    os.makedirs(BUILD, exist_ok=True)
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
      subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, DEPS[0]])
    _lib = ctypes.CDLL(LIB)
    _lib.emu_last_error.restype = ctypes.c_char_p
    _lib.emu_trainer_create.restype = ctypes.c_void_p
    _lib.emu_trainer_create.argtypes = [ctypes.POINTER(TT), ctypes.c_int32, ctypes.c_int32]
    _lib.emu_trainer_destroy.argtypes = [ctypes.c_void_p]
    _lib.emu_forward_backward.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int32] * 2 + [ctypes.c_void_p] * 3
    _lib.emu_activation.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p),
                                    ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32)]
    f = ctypes.c_float
    _lib.emu_adam_step.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64, ctypes.c_int32] + [f] * 7
  return _lib


def _p(t):
  return None if t is None else ctypes.c_void_p(t.data_ptr())


def branch_from_activations(fetch, B):
  """{unit: (pass-through mask, saturated mask)} for `oracle.restatement` from the
  post-activation outputs of a training step.  `fetch(index)` returns the [rows, channels]
  activation of unit `index` (0..51 conv+BN units, NHWC rows; 52..54 merger layers)."""
  branch = {}
  for i in range(55):
    a = fetch(i)
    if i < 52:
      hw = a.shape[0] // B
      h = int(round(hw**0.5))
      a = a.view(B, h, h, a.shape[1]).permute(0, 3, 1, 2)
    branch[i] = ((a > 0) & (a < 6) if i < 52 else a > 0, a >= 6)
  return branch


class EmuTrainer:
  """Host-memory twin of oatomobile_b200.train.Trainer's native half."""

  def __init__(self, state_dict, kind):
    self.kind = kind
    self.params = {k: v.detach().clone().float().contiguous() for k, v in state_dict.items()
                   if v.is_floating_point()}
    self.grads = {k: torch.full_like(v, float("nan")) for k, v in self.params.items()
                  if "running_" not in k}
    arr = []
    for k, v in self.params.items():
      shape = (ctypes.c_int64 * 4)(*(list(v.shape) + [0] * (4 - v.dim())))
      g = self.grads[k].data_ptr() if k in self.grads else None
      arr.append(TT(k.encode(), v.data_ptr(), g, v.dim(), shape))
    self._arr = (TT * len(arr))(*arr)
    self._ptr = lib().emu_trainer_create(self._arr, len(arr), 0 if kind == "dim" else 1)
    if not self._ptr:
      raise RuntimeError(lib().emu_last_error().decode())
    self.m = {k: torch.zeros_like(v) for k, v in self.grads.items()}
    self.v = {k: torch.zeros_like(v) for k, v in self.grads.items()}
    self.steps = 0

  def __del__(self):
    if getattr(self, "_ptr", None):
      lib().emu_trainer_destroy(self._ptr)
      self._ptr = None

  def forward_backward(self, visual, scalars, target, mask=None):
    visual, scalars, target = (t.detach().float().contiguous() for t in (visual, scalars, target))
    mask = None if mask is None else mask.detach().float().contiguous()
    B, T = target.shape[0], target.shape[1]
    loss, z, pred = torch.zeros(1), torch.zeros(B, 64), torch.zeros(B, T, 2)
    rc = lib().emu_forward_backward(self._ptr, _p(visual), _p(scalars), _p(target), _p(mask), B, T,
                                    _p(loss), _p(z), _p(pred))
    if rc:
      raise RuntimeError(lib().emu_last_error().decode())
    return loss[0], (z if self.kind == "dim" else pred)

  def activation(self, index):
    data, rows, ch = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int32()
    if lib().emu_activation(self._ptr, index, ctypes.byref(data), ctypes.byref(rows), ctypes.byref(ch)):
      raise RuntimeError("emu_activation(%d) failed" % index)
    n = rows.value * ch.value
    buf = (ctypes.c_float * n).from_address(data.value)
    return torch.frombuffer(buf, dtype=torch.float32).clone().view(rows.value, ch.value)

  def branch(self, B):
    return branch_from_activations(self.activation, B)

  def adam(self, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0, clip_norm=0.0):
    """Per-tensor Adam (clip_norm needs the global norm → only valid with a flat buffer;
    the tests exercise clipping on a single tensor)."""
    self.steps += 1
    for k, g in self.grads.items():
      lib().emu_adam_step(_p(self.params[k]), _p(g), _p(self.m[k]), _p(self.v[k]), g.numel(),
                          self.steps, lr, betas[0], betas[1], eps, weight_decay, grad_scale, clip_norm)
