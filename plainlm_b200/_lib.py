"""ctypes binding of libplainlm_b200.so (the C ABI declared in include/plainlm_b200.h).

There is no fallback: if the shared library is missing, or a call returns a negative status, this raises.
"""

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libplainlm_b200.so')
CSRC_DIR = os.path.join(_HERE, 'csrc')

PLM_OK = 0
EPI_BF16, EPI_BF16_ROPE, EPI_F32, EPI_RESID_F32, EPI_ATOMIC_F32, EPI_BF16_SWIGLU, EPI_BF16_CE, EPI_BF16_GLU_BWD = range(8)
SUMSQ_WORKSPACE = 1024
ACT_SILU, ACT_RELU2 = 0, 1

c_void_p, c_int32, c_int64, c_float = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float


class GemmArgs(ctypes.Structure):
  _fields_ = [
    ('A', c_void_p), ('B', c_void_p), ('C', c_void_p), ('R', c_void_p), ('rope_table', c_void_p),
    ('M', c_int64), ('N', c_int64), ('K', c_int64),
    ('lda', c_int64), ('ldb', c_int64), ('ldc', c_int64),
    ('a_kmajor', c_int32), ('b_kmajor', c_int32), ('epilogue', c_int32), ('splits', c_int32),
    ('rope_cols', c_int32), ('rope_T', c_int32), ('head_dim', c_int32),
    ('C2', c_void_p), ('ldc2', c_int64),
    ('ce_targets', c_void_p), ('ce_partial', c_void_p), ('ce_tgt_logit', c_void_p),
  ]  # fmt: skip


# name -> (restype, argtypes); mirrors include/plainlm_b200.h one to one
_P, _I32, _I64, _F = c_void_p, c_int32, c_int64, c_float
SIGNATURES = {
  'plm_abi_version': (c_int32, []),
  'plm_last_error': (ctypes.c_char_p, []),
  'plm_device_check': (c_int32, []),
  'plm_gemm_bf16': (c_int32, [ctypes.POINTER(GemmArgs), _P]),
  'plm_gemm_set_tuning': (c_int32, [_I32, _I32, _I32, _I32, _I32]),
  'plm_attn_fwd': (c_int32, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _P]),
  'plm_attn_fwd_variant': (c_int32, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P]),
  'plm_attn_fwd_v1': (c_int32, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _P]),
  'plm_attn_bwd': (c_int32, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P]),
  'plm_attn_bwd_variant': (c_int32, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P]),
  'plm_rope_qk': (c_int32, [_P, _P, _I64, _I32, _I32, _I32, _I32, _P]),
  'plm_rmsnorm_fwd': (c_int32, [_P, _P, _P, _P, _I64, _I32, _F, _P]),
  'plm_rmsnorm_bwd_blocks': (c_int32, [_I64]),
  'plm_rmsnorm_bwd': (c_int32, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
  'plm_colsum_accum': (c_int32, [_P, _P, _I32, _I32, _P]),
  'plm_colsum_accum_batched': (c_int32, [_P, _P, _I32, _I32, _I32, _P]),
  'plm_swiglu_fwd': (c_int32, [_P, _P, _I64, _I32, _P]),
  'plm_swiglu_bwd': (c_int32, [_P, _P, _P, _I64, _I32, _P]),
  'plm_act_fwd': (c_int32, [_P, _P, _I64, _I32, _P]),
  'plm_act_bwd': (c_int32, [_P, _P, _P, _I64, _I32, _P]),
  'plm_embed_fwd': (c_int32, [_P, _P, _P, _I64, _I32, _I64, _P]),
  'plm_embed_bwd': (c_int32, [_P, _P, _P, _I64, _I32, _I64, _P]),
  'plm_ce_fwd_bwd': (c_int32, [_P, _P, _P, _P, _P, _I64, _I32, _I64, _F, _I32, _P]),
  'plm_lmhead_ce_tiles': (c_int32, [_I64]),
  'plm_lmhead_ce_fwd': (c_int32, [_P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I64, _I64, _P]),
  'plm_ce_grad': (c_int32, [_P, _P, _P, _P, _I64, _I32, _I64, _F, _P]),
  'plm_sumsq': (c_int32, [_P, _I64, _P, _P, _I32, _P]),
  'plm_unpack_sumsq': (c_int32, [_P, _P, _I64, _F, _P, _P, _P]),
  'plm_adamw_step': (c_int32, [_P, _P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _F, _F, _P, _F, _P]),
  'plm_signsgd_step': (c_int32, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _I32, _P, _F, _P]),
  'plm_nadamw_step': (c_int32, [_P, _P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _F, _F, _F, _P, _F, _P]),
  'plm_sgd_step': (c_int32, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _I32, _P, _F, _P]),
  'plm_cast_f32_bf16': (c_int32, [_P, _P, _I64, _F, _P]),
  'plm_cast_bf16_f32': (c_int32, [_P, _P, _I64, _F, _P]),
  'plm_seg_start_from_lengths': (c_int32, [_P, _P, _P, _I32, _I32, _P]),
}  # fmt: skip

_lib = None


def build(verbose=False):
  """Compile libplainlm_b200.so for sm_100a with nvcc (in-tree, via csrc/Makefile)."""
  cmd = ['make', '-C', CSRC_DIR, '-j', str(os.cpu_count() or 4)]
  res = subprocess.run(cmd, capture_output=True, text=True)
  if verbose or res.returncode != 0:
    print(res.stdout)
    print(res.stderr)
  if res.returncode != 0:
    raise RuntimeError('building libplainlm_b200.so failed')
  return LIB_PATH


def load():
  """Load the shared library and attach signatures. Raises if it is missing: there is no CPU path."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise RuntimeError(
      f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
      '(plainlm_b200 has no fallback path)'
    )
  lib = ctypes.CDLL(LIB_PATH)
  for name, (restype, argtypes) in SIGNATURES.items():
    fn = getattr(lib, name)  # AttributeError if the symbol is missing
    fn.restype = restype
    fn.argtypes = argtypes
  if lib.plm_abi_version() != 5:
    raise RuntimeError('libplainlm_b200.so ABI version mismatch')
  _lib = lib
  return lib


def gemm_tuning(bn=0, raster=-1, cluster=0, pair=1, debug=0):
  """Diagnostics: process-global overrides of plm_gemm_bf16's automatic tile choices (no arguments = defaults)."""
  check(load().plm_gemm_set_tuning(bn, raster, cluster, pair, debug), 'plm_gemm_set_tuning')


class PlmError(RuntimeError):
  pass


def check(status, what):
  if status != PLM_OK:
    msg = load().plm_last_error().decode('utf-8', 'replace')
    raise PlmError(f'{what} failed with status {status}: {msg}')
