"""
ctypes binding of the C ABI in include/gsr_b200.h (csrc/libgsr_b200.so).

There is NO fallback: if the CUDA library cannot be loaded, or a tensor is not a contiguous float32 / int32
CUDA tensor, the call raises.  PyTorch only owns the memory and the stream.
"""
import ctypes as C
import os

import torch

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

GSR_NSETS = 3


class GridDesc(C.Structure):
	_fields_ = [('D', C.c_int32), ('dims', C.c_int32 * 3), ('lo', C.c_float * 3), ('hi', C.c_float * 3),
				('grid_scale', C.c_float), ('tau', C.c_float), ('grid_scale_dev', C.c_void_p)]


class LossCfg(C.Structure):
	_fields_ = [('w_val', C.c_float), ('w_boundary', C.c_float), ('w_grad', C.c_float), ('w_vor', C.c_float), ('w_hel', C.c_float), ('w_div', C.c_float),
				('Q_norm', C.c_int64),
				('ref_val', C.c_void_p), ('normals', C.c_void_p), ('normal_ref', C.c_void_p), ('ref_grad', C.c_void_p),
				('ref_vor', C.c_void_p), ('ref_hel', C.c_void_p), ('stop_gradient', C.c_void_p), ('sample_grid_scale_dev', C.c_void_p), ('loss_partials', C.c_void_p)]


class StepCfg(C.Structure):
	_fields_ = [('D', C.c_int32), ('lr', C.c_float * 4), ('beta1', C.c_float), ('beta2', C.c_float), ('eps', C.c_float),
				('sched_factor', C.c_float), ('sched_threshold', C.c_float), ('sched_eps', C.c_float), ('sched_min_lr', C.c_float),
				('sched_patience', C.c_int32),
				('w_aniso', C.c_float), ('w_vol', C.c_float), ('w_valreg', C.c_float), ('w_dpos', C.c_float),
				('aniso_ratio', C.c_float), ('pcgrad', C.c_int32),
				('grid_coef', C.c_double), ('min_grid_scale', C.c_double), ('grid_scale_tau0', C.c_double),
				('keep_clock', C.c_int32), ('sample_gs_slots', C.c_void_p), ('sample_gs_margin', C.c_float), ('grid_scale_out', C.c_void_p)]


class LossSrc(C.Structure):
	_fields_ = [('partials', C.c_void_p), ('nblocks', C.c_int32), ('w', C.c_float * 8)]


# indices into the device-resident optimiser state (include/gsr_b200.h)
STATE_SCALARS = 64
ST_T, ST_BEST, ST_BAD, ST_LR, ST_GRID_SCALE, ST_MIN_S, ST_LOSS_TOT, ST_L_ANISO, ST_L_VOL, ST_L_VALREG, ST_L_DPOS = 0, 1, 2, 3, 7, 8, 9, 10, 11, 12, 13
ST_CLOCK = 22
ST_SGS_ERR = 23


class GsrError(RuntimeError):
	pass


def lib():
	"""Load (building first if the sources are newer) the CUDA library.  Raises if that is impossible."""
	global _lib
	if _lib is None:
		so = _build.SO
		if _build.needs_build():
			so = _build.build()
		_lib = C.CDLL(so)
		for name in ('gsr_build_grid_ws_bytes', 'gsr_bin_samples_ws_bytes', 'gsr_backward_ws_bytes', 'gsr_step_ws_bytes', 'gsr_step_state_floats'):
			if hasattr(_lib, name):
				getattr(_lib, name).restype = C.c_size_t
		_lib.gsr_padded_cells.restype = C.c_int64
		_lib.gsr_tile_slots.restype = C.c_int64
		_lib.gsr_loss_blocks.restype = C.c_int64
		_lib.gsr_launch_count.restype = C.c_uint64
		_lib.gsr_version.restype = C.c_char_p
	return _lib


def check(rc, what):
	if rc != 0:
		msg = {-1: 'invalid argument', -2: 'workspace too small'}.get(rc, f'CUDA error {rc}')
		raise GsrError(f'{what}: {msg}')


def ptr(t, dtype=torch.float32, allow_none=False, name='tensor', align16=False):
	"""device pointer of a contiguous CUDA tensor of the given dtype (no silent copies, no CPU fallback)"""
	if t is None:
		if allow_none:
			return None
		raise GsrError(f'{name} is None')
	if not isinstance(t, torch.Tensor) or not t.is_cuda:
		raise GsrError(f'{name} must be a CUDA tensor (the engine has no CPU path)')
	if t.dtype != dtype:
		raise GsrError(f'{name} must be {dtype}, got {t.dtype}')
	if not t.is_contiguous():
		raise GsrError(f'{name} must be contiguous')
	if align16 and t.numel() and t.data_ptr() % 16:
		raise GsrError(f'{name} must be 16-byte aligned')
	return C.c_void_p(t.data_ptr()) if t.numel() else C.c_void_p(0)


def stream():
	return C.c_void_p(torch.cuda.current_stream().cuda_stream)
