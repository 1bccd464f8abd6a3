"""TEST TOOLING — builds and drives tests/emu/fused_emu.cpp: the fused encoder-front CTA
bodies of oatomobile_b200/csrc/fused_body.h executed on the host (phase = loop over thread
ids).  Used by the CPU unit tests only; never imported by the oatomobile_b200 package."""
import ctypes
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libfused_emu.so")
DEPS = [os.path.join(HERE, "fused_emu.cpp"),
        os.path.join(ROOT, "oatomobile_b200", "csrc", "fused_body.h")]

_lib = None


def lib():
  global _lib
  if _lib is None:
    os.makedirs(BUILD, exist_ok=True)
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
      subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, DEPS[0]])
    _lib = ctypes.CDLL(LIB)
    p, i = ctypes.c_void_p, ctypes.c_int
    _lib.emu_expand_dw.argtypes = [i, p, i, p, p, p, p, p, i, i]
    _lib.emu_expand_dw_tc.argtypes = [i, p, i, p, p, p, p, p, i, i, i]
    _lib.emu_front.argtypes = [p, i, i, p, p, p, p, p, p, p, i, i]
    _lib.emu_dw_project.argtypes = [p, i, p, p, p, p, p, i]
  return _lib


def _p(t):
  return ctypes.c_void_p(t.data_ptr())


def fold_bn(sd, bn):
  """Eval-mode BatchNorm as (scale, shift) in float64 — what api.cu's Packer::bn does."""
  g, b = sd[bn + ".weight"].double(), sd[bn + ".bias"].double()
  m, v = sd[bn + ".running_mean"].double(), sd[bn + ".running_var"].double()
  scale = g / torch.sqrt(v + 1e-5)
  return scale, b - m * scale


def pack_pw(sd, conv, bn):
  """1x1 conv [N,K,1,1] + BN -> w [K][N], b [N] (api.cu pack_pw)."""
  scale, shift = fold_bn(sd, bn)
  w = sd[conv + ".weight"].double()[:, :, 0, 0] * scale[:, None]
  return w.t().contiguous().float(), shift.float().contiguous()


def pack_dw(sd, conv, bn):
  """depthwise 3x3 [C,1,3,3] + BN -> w [9][C], b [C] (api.cu pack_dw)."""
  scale, shift = fold_bn(sd, bn)
  w = sd[conv + ".weight"].double()[:, 0].reshape(-1, 9) * scale[:, None]
  return w.t().contiguous().float(), shift.float().contiguous()


def pack_stem(sd, conv, bn):
  """3x3 conv [32,C,3,3] + BN -> w [(kh*3+kw)*C + c][32], b [32] (api.cu, stem)."""
  scale, shift = fold_bn(sd, bn)
  w = sd[conv + ".weight"].double() * scale[:, None, None, None]  # [32,C,3,3]
  return w.permute(2, 3, 1, 0).reshape(-1, 32).contiguous().float(), shift.float().contiguous()


def expand_dw(idx, sd, x_nhwc, splits=2, threads=256, prefix="_encoder._model.features.", tc=False,
              ctas=3):
  """features.<idx> expand + depthwise on the host.  x_nhwc [B,H,H,cin] -> [B,Ho,Ho,hid].
  tc=True runs the pipelined tensor-core body (GEMM by plain host loops, `ctas` persistent
  executors sharing the work units)."""
  p = prefix + "%d.conv" % idx
  we, be = pack_pw(sd, p + ".0.0", p + ".0.1")
  wd, bd = pack_dw(sd, p + ".1.0", p + ".1.1")
  B, H = x_nhwc.shape[0], x_nhwc.shape[1]
  stride = {2: 2, 3: 1, 4: 2}[idx]
  Ho = (H - 1) // stride + 1
  x = x_nhwc.contiguous().float()
  out = torch.full((B, Ho, Ho, we.shape[1]), float("nan"))
  if tc:
    rc = lib().emu_expand_dw_tc(idx, _p(x), B, _p(we), _p(be), _p(wd), _p(bd), _p(out), splits,
                                threads, ctas)
  else:
    rc = lib().emu_expand_dw(idx, _p(x), B, _p(we), _p(be), _p(wd), _p(bd), _p(out), splits, threads)
  assert rc == 0
  return out


def front(sd, visual, splits=2, threads=256, prefix="_encoder._model.features."):
  """features.0 + features.1 on the host.  visual [B,C,100,100] -> [B,50,50,16] (NHWC)."""
  ws, bs = pack_stem(sd, prefix + "0.0", prefix + "0.1")
  wd, bd = pack_dw(sd, prefix + "1.conv.0.0", prefix + "1.conv.0.1")
  wp, bp = pack_pw(sd, prefix + "1.conv.1", prefix + "1.conv.2")
  B, C = visual.shape[0], visual.shape[1]
  v = visual.contiguous().float()
  out = torch.full((B, 50, 50, 16), float("nan"))
  rc = lib().emu_front(_p(v), B, C, _p(ws), _p(bs), _p(wd), _p(bd), _p(wp), _p(bp), _p(out), splits, threads)
  assert rc == 0
  return out


def dw_project(sd, x_nhwc, threads=128, prefix="_encoder._model.features."):
  """features.1 (depthwise + project) on the host.  x_nhwc [B,50,50,32] -> [B,50,50,16]."""
  wd, bd = pack_dw(sd, prefix + "1.conv.0.0", prefix + "1.conv.0.1")
  wp, bp = pack_pw(sd, prefix + "1.conv.1", prefix + "1.conv.2")
  x = x_nhwc.contiguous().float()
  B = x.shape[0]
  out = torch.full((B, 50, 50, 16), float("nan"))
  rc = lib().emu_dw_project(_p(x), B, _p(wd), _p(bd), _p(wp), _p(bp), _p(out), threads)
  assert rc == 0
  return out
