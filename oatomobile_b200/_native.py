"""ctypes binding of `liboat_b200.so` (the C-ABI declared in include/oat_b200.h).

The library is the product: there is no PyTorch/CPU fallback.  If the shared
object is missing, or no CUDA device is present, every entry point raises.
"""
import ctypes
import os
import threading
from typing import Dict, Mapping, Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OAT_B200_LIB", os.path.join(_HERE, "liboat_b200.so"))

KIND_DIM, KIND_CIL, KIND_FLOW, KIND_ENCODER = 0, 1, 2, 3
ALGORITHMS = {"WCM": 0, "BCM": 1, "MA": 2}

# Every symbol include/oat_b200.h declares (checked by tests/test_abi.py).
SYMBOLS = (
    "oat_last_error", "oat_abi_version", "oat_model_create", "oat_model_destroy",
    "oat_model_in_channels", "oat_ensemble_create", "oat_ensemble_destroy",
    "oat_ensemble_reserve", "oat_transform_visual", "oat_encode", "oat_flow_forward",
    "oat_flow_inverse", "oat_rip_sample_score", "oat_rip_aggregate", "oat_cil_rollout",
    "oat_launch_count", "oat_ensemble_set_pw_impl", "oat_debug_tc_gemm",
    "oat_set_flow_impl", "oat_plan", "oat_plan_workspace_floats", "oat_goal_likelihood",
    "oat_transform_visual_hwc", "oat_lidar_bev", "oat_trainer_create", "oat_trainer_destroy",
    "oat_train_forward_backward", "oat_adam_step", "oat_trainer_activation",
    "oat_ensemble_set_fusion", "oat_ensemble_get_fusion", "oat_debug_encoder_prefix",
    "oat_ensemble_set_fusion_tc", "oat_profile_begin", "oat_profile_end",
    "oat_encode_features", "oat_mlp_forward",
)


class NativeLibraryError(RuntimeError):
  """The CUDA library is missing or a native call failed."""


class OatTensor(ctypes.Structure):
  _fields_ = [("name", ctypes.c_char_p), ("h_data", ctypes.c_void_p),
              ("ndim", ctypes.c_int32), ("shape", ctypes.c_int64 * 4)]


class OatTrainTensor(ctypes.Structure):
  _fields_ = [("name", ctypes.c_char_p), ("param", ctypes.c_void_p), ("grad", ctypes.c_void_p),
              ("ndim", ctypes.c_int32), ("shape", ctypes.c_int64 * 4)]


_lib = None
_lock = threading.Lock()


def lib() -> ctypes.CDLL:
  """Loads the in-tree shared object; raises loudly when it is absent."""
  global _lib
  if _lib is not None:
    return _lib
  with _lock:
    if _lib is not None:
      return _lib
    if not os.path.exists(LIB_PATH):
      raise NativeLibraryError(
          "oatomobile_b200: %s not found. Build it with `python -m oatomobile_b200.build` "
          "(nvcc, sm_100a). There is no CPU/PyTorch fallback for this path." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    c_int, c_i32, c_i64, c_f, vp = (ctypes.c_int, ctypes.c_int32, ctypes.c_int64,
                                    ctypes.c_float, ctypes.c_void_p)
    L.oat_last_error.restype = ctypes.c_char_p
    L.oat_abi_version.restype = c_int
    L.oat_launch_count.restype = c_i64
    L.oat_model_create.argtypes = [ctypes.POINTER(OatTensor), c_i32, c_i32, c_i32,
                                   ctypes.POINTER(vp)]
    L.oat_model_destroy.argtypes = [vp]
    L.oat_model_in_channels.argtypes = [vp]
    L.oat_ensemble_create.argtypes = [ctypes.POINTER(vp), c_i32, ctypes.POINTER(vp)]
    L.oat_ensemble_destroy.argtypes = [vp]
    L.oat_ensemble_reserve.argtypes = [vp, c_i32]
    L.oat_transform_visual.argtypes = [vp, c_i32, c_i32, c_i32, c_i32, vp, vp]
    L.oat_encode.argtypes = [vp, vp, vp, c_i32, vp, vp]
    L.oat_transform_visual_hwc.argtypes = [vp, c_i32, c_i32, c_i32, c_i32, vp, vp]
    L.oat_flow_forward.argtypes = [vp, vp, vp, c_i64, c_i32, c_i32, vp, vp, vp]
    L.oat_flow_inverse.argtypes = [vp, vp, vp, c_i64, c_i32, c_i32, vp, vp, vp, vp]
    L.oat_rip_sample_score.argtypes = [vp, c_i32, vp, vp, vp, c_i32, c_f, c_i32, c_i32, c_i32,
                                       vp, vp, vp]
    L.oat_rip_aggregate.argtypes = [vp, c_i32, c_i32, c_i32, c_i32, vp, c_i32, vp, vp, vp, vp,
                                    vp]
    L.oat_cil_rollout.argtypes = [vp, vp, c_i32, c_i32, vp, vp]
    L.oat_ensemble_set_pw_impl.argtypes = [vp, c_i32]
    L.oat_ensemble_set_fusion.argtypes = [vp, c_i32]
    L.oat_ensemble_get_fusion.argtypes = [vp]
    L.oat_ensemble_set_fusion_tc.argtypes = [vp, c_i32]
    L.oat_debug_encoder_prefix.argtypes = [vp, vp, c_i32, c_i32, vp, vp]
    L.oat_set_flow_impl.argtypes = [c_i32]
    L.oat_encode_features.argtypes = [vp, vp, c_i32, vp, vp]
    L.oat_mlp_forward.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(c_i32), c_i32,
                                  c_i32, vp, c_i32, vp, vp]
    L.oat_profile_begin.argtypes = [vp]
    L.oat_profile_end.argtypes = [ctypes.c_char_p, c_i64]
    L.oat_plan.argtypes = [ctypes.POINTER(vp), c_i32, c_i32, vp, vp, c_i32, c_f, c_i32, c_i32, c_i32,
                           c_f, vp, vp, vp, vp, c_i64, vp, vp]
    L.oat_plan_workspace_floats.argtypes = [c_i32, c_i32, c_i32]
    L.oat_lidar_bev.argtypes = [vp, c_i64, c_i32, c_i32, c_i32, vp, vp, vp]
    L.oat_goal_likelihood.argtypes = [vp, vp, c_i32, c_i32, c_f, vp, vp, vp]
    L.oat_debug_tc_gemm.argtypes = [vp, vp, vp, vp, vp, c_i32, c_i32, c_i32, c_i32, c_i32, vp]
    L.oat_trainer_create.argtypes = [ctypes.POINTER(OatTrainTensor), c_i32, c_i32, c_i32, vp, c_i64,
                                     ctypes.POINTER(vp)]
    L.oat_trainer_destroy.argtypes = [vp]
    L.oat_train_forward_backward.argtypes = [vp, vp, vp, vp, vp, c_i32, c_i32, vp, vp, vp, vp]
    L.oat_trainer_activation.argtypes = [vp, c_i32, vp, ctypes.POINTER(c_i64),
                                         ctypes.POINTER(c_i32), vp]
    L.oat_adam_step.argtypes = [vp, vp, vp, vp, c_i64, c_i32, c_f, c_f, c_f, c_f, c_f, c_f, c_f,
                                vp, vp]
    for name in SYMBOLS:
      fn = getattr(L, name, None)
      if fn is not None and name not in ("oat_last_error", "oat_launch_count",
                                         "oat_plan_workspace_floats"):
        fn.restype = c_int
    L.oat_plan_workspace_floats.restype = c_i64
    _lib = L
    if os.environ.get("OAT_FLOW_IMPL"):  # A/B switch: simt | tcgen05 | tcgen05x2
      L.oat_set_flow_impl({"simt": 0, "tcgen05": 1, "tcgen05x2": 2}[os.environ["OAT_FLOW_IMPL"]])
    return L


def check(rc: int) -> None:
  if rc != 0:
    raise NativeLibraryError(lib().oat_last_error().decode() or "native call failed")


_default_pw_impl = "tcgen05"


def set_default_pw_impl(impl: str) -> None:
  """Kernel family new ensembles use for the pointwise convolutions ("tcgen05"|"simt")."""
  global _default_pw_impl
  assert impl in ("tcgen05", "simt")
  _default_pw_impl = impl


_default_fusion = None  # None: the library's default (OAT_FUSE_DEFAULT / env OAT_FUSE)


def set_default_fusion(mask) -> None:
  """Fusion mask (see `oat_ensemble_set_fusion`) for ensembles created from now on;
  None restores the library default."""
  global _default_fusion
  if mask is not None and not 0 <= int(mask) <= 63:
    raise ValueError("fusion mask must be in [0, 63]")
  _default_fusion = None if mask is None else int(mask)


def set_flow_impl(impl: str) -> None:
  """"tcgen05" (default) or "simt": kernel family of the autoregressive flow."""
  check(lib().oat_set_flow_impl({"simt": 0, "tcgen05": 1, "tcgen05x2": 2}[impl]))


def launch_count() -> int:
  return int(lib().oat_launch_count())


def stream_ptr(device: torch.device) -> int:
  return torch.cuda.current_stream(device).cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
  return None if t is None else t.data_ptr()


def require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
  """The kernels read contiguous float32 device memory; anything else is an error
  (CPU tensors in particular: this path has no CPU implementation)."""
  if not isinstance(t, torch.Tensor):
    raise TypeError("`%s` must be a torch.Tensor" % name)
  if not t.is_cuda:
    raise NativeLibraryError(
        "`%s` lives on %s: oatomobile_b200 runs on CUDA (sm_100a) only, there is no CPU "
        "fallback. Move the model and its inputs to a GPU." % (name, t.device))
  if t.dtype != torch.float32:
    t = t.float()
  return t.contiguous()


class ModelHandle:
  """Owns one `OatModel*` built from a reference-format state_dict."""

  def __init__(self, state_dict: Mapping[str, torch.Tensor], kind: int, device: torch.device):
    L = lib()
    if device.type != "cuda":
      raise NativeLibraryError("oatomobile_b200 models run on CUDA devices only (got %s)" % device)
    index = device.index if device.index is not None else torch.cuda.current_device()
    keep, arr = [], []
    for name, t in state_dict.items():
      if not torch.is_floating_point(t):
        continue  # num_batches_tracked (int64) is unused in eval mode
      h = t.detach().to("cpu", torch.float32).contiguous()
      keep.append(h)
      shape = (ctypes.c_int64 * 4)(*(list(h.shape) + [0] * (4 - h.dim())))
      arr.append(OatTensor(name.encode(), h.data_ptr(), h.dim(), shape))
    c_arr = (OatTensor * len(arr))(*arr)
    out = ctypes.c_void_p()
    check(L.oat_model_create(c_arr, len(arr), kind, index, ctypes.byref(out)))
    self.ptr = out
    self.kind = kind
    self.device = torch.device("cuda", index)

  def __deepcopy__(self, memo):  # device handles are never shared between copies
    return None

  def __del__(self):
    try:
      if getattr(self, "ptr", None):
        lib().oat_model_destroy(self.ptr)
        self.ptr = None
    except Exception:
      pass


class EnsembleHandle:
  """Owns one `OatEnsemble*` over E model handles living on the same GPU."""

  _generation = 0

  def __init__(self, models: Sequence[ModelHandle]):
    L = lib()
    self.models = list(models)  # keep the model handles alive
    arr = (ctypes.c_void_p * len(models))(*[m.ptr.value for m in models])
    out = ctypes.c_void_p()
    check(L.oat_ensemble_create(arr, len(models), ctypes.byref(out)))
    self.ptr = out
    self.device = models[0].device
    self.in_channels = int(L.oat_model_in_channels(models[0].ptr))
    self.scalars = {KIND_CIL: 6, KIND_ENCODER: 0}.get(models[0].kind, 5)  # merger inputs besides the 128 features
    EnsembleHandle._generation += 1
    self.generation = EnsembleHandle._generation  # never reused (unlike id()): keys CUDA graphs
    if _default_pw_impl != "tcgen05":
      self.set_pw_impl(_default_pw_impl)
    if _default_fusion is not None:
      self.set_fusion(_default_fusion)

  def __len__(self):
    return len(self.models)

  def set_pw_impl(self, impl: str) -> None:
    """"tcgen05" (3xTF32 tensor-core GEMMs, default) or "simt" (FP32 FFMA GEMMs)."""
    check(lib().oat_ensemble_set_pw_impl(self.ptr, {"simt": 0, "tcgen05": 1}[impl]))

  def set_fusion(self, mask: int) -> None:
    """Bit 0: features.0+1 as one kernel; bits 1-3: expand+depthwise of features.2-4 fused;
    bit 4: depthwise+project of features.1 fused; bit 5: expand+depthwise of features.5-17 inside
    the tcgen05 GEMM (depthwise epilogue)."""
    check(lib().oat_ensemble_set_fusion(self.ptr, int(mask)))

  def fusion(self) -> int:
    return int(lib().oat_ensemble_get_fusion(self.ptr))

  def set_fusion_tc(self, mode: int) -> None:
    """Fused expand GEMM: 0 FP32 FMA, 1 auto (default), 2 tcgen05 for every fused block."""
    check(lib().oat_ensemble_set_fusion_tc(self.ptr, int(mode)))

  def __deepcopy__(self, memo):
    return None

  def __del__(self):
    try:
      if getattr(self, "ptr", None):
        lib().oat_ensemble_destroy(self.ptr)
        self.ptr = None
    except Exception:
      pass


def profile_begin(device) -> None:
  """Opens a per-kernel-family device timing (include/oat_b200.h: oat_profile_begin)."""
  with torch.cuda.device(device):
    check(lib().oat_profile_begin(stream_ptr(device)))


def profile_end() -> dict:
  """Closes it: {"family": {"ms": total, "launches": n}, ...}."""
  import json
  buf = ctypes.create_string_buffer(1 << 16)
  check(lib().oat_profile_end(buf, len(buf)))
  return json.loads(buf.value.decode())
