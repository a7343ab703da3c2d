"""Thin functional layer over the C ABI: builds ``azp_pair_args`` from device tensors and
enqueues one force evaluation on the current CUDA stream. Used by ``pair.Pair.compute`` and by
the multi-GPU particle-slice scheduler (``slices.py``)."""

import ctypes

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return None if t is None else t.data_ptr()


def scalar_bits(t):
    return 32 if t.dtype == torch.float32 else 64


def fill_args(*, box, pos, n_neigh, nlist, head_list, rcutsq, ntypes, force, n_rows=None,
              ronsq=None, virial=None, torque=None, vel=None, orientation=None, tag=None,
              shift_mode=0, compute_virial=False, block_size=0, threads_per_particle=0,
              seed=0, timestep=0, dt=0.0, kT=0.0, row_offset=0, row_ids=None,
              size_neigh_list=None, n_max=0, virial_pitch=None):
    a = _lib.AzpPairArgs()
    a.d_force = _ptr(force)
    a.d_virial = _ptr(virial)
    a.d_torque = _ptr(torque)
    a.virial_pitch = 0 if virial is None else int(virial.shape[-1] if virial_pitch is None else virial_pitch)
    a.d_pos = _ptr(pos)
    a.d_vel = _ptr(vel)
    a.d_orientation = _ptr(orientation)
    a.d_tag = _ptr(tag)
    a.d_n_neigh = _ptr(n_neigh)
    a.d_nlist = _ptr(nlist)
    a.d_head_list = _ptr(head_list)
    a.size_neigh_list = int(nlist.numel() if size_neigh_list is None else size_neigh_list)
    a.d_rcutsq = _ptr(rcutsq)
    a.d_ronsq = _ptr(ronsq)
    a.box = box.to_c()
    a.N = int(n_neigh.numel() if n_rows is None else n_rows)
    a.ntypes = int(ntypes)
    a.shift_mode = int(shift_mode)
    a.compute_virial = int(bool(compute_virial))
    a.block_size = int(block_size)
    a.threads_per_particle = int(threads_per_particle)
    a.seed = int(seed) & 0xFFFF
    a.n_max = int(n_max)
    a.timestep = int(timestep)
    a.deltaT = float(dt)
    a.T = float(kT)
    a.row_offset = int(row_offset)
    if row_ids is not None:
        a.d_row_ids = _ptr(row_ids)
        a.n_row_ids = int(row_ids.numel())
    return a


def launch(family, evaluator, bits, args, d_params, stream=None):
    """Enqueue one evaluation. ``stream``: raw cudaStream_t (int) or None = torch current."""
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    st = ctypes.c_void_p(stream)
    sfx = "_f%d" % bits
    if family == _lib.FAMILY_PAIR:
        rc = getattr(_lib.lib, "azp_pair_forces" + sfx)(evaluator, ctypes.byref(args), d_params, st)
    elif family == _lib.FAMILY_DPD:
        rc = getattr(_lib.lib, "azp_dpd_forces" + sfx)(evaluator, ctypes.byref(args), d_params, st)
    elif family == _lib.FAMILY_ANISO:
        rc = getattr(_lib.lib, "azp_aniso_forces" + sfx)(evaluator, ctypes.byref(args), d_params, None, st)
    else:
        raise ValueError("unknown kernel family")
    _lib.check(rc, "pair-force launch")


def launch_fused(ev_a, args_a, d_params_a, ev_b, args_b, d_params_b, bits, stream=None):
    """Two isotropic potentials over one sweep of the list (``azp_pair_forces_fused_*``)."""
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    fn = getattr(_lib.lib, "azp_pair_forces_fused_f%d" % bits)
    rc = fn(ev_a, ctypes.byref(args_a), d_params_a, ev_b, ctypes.byref(args_b), d_params_b,
            ctypes.c_void_p(stream))
    _lib.check(rc, "fused pair-force launch")


def autotune(family, evaluator, bits, args, d_params, stream=None):
    """(block_size, threads_per_particle, ms) of the fastest launch shape for these arguments."""
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    bb, bt, ms = ctypes.c_uint32(0), ctypes.c_uint32(0), ctypes.c_float(0)
    rc = _lib.lib.azp_autotune(family, evaluator, bits, ctypes.byref(args), d_params,
                               ctypes.c_void_p(stream), ctypes.byref(bb), ctypes.byref(bt),
                               ctypes.byref(ms))
    _lib.check(rc, "autotune")
    return int(bb.value), int(bt.value), float(ms.value)


def param_size(evaluator, bits):
    return int(_lib.lib.azp_param_size(evaluator, bits))


def pack_params(evaluator, bits, fields):
    """``param_type`` bytes (numpy uint8) for one type pair from its double fields."""
    f = np.ascontiguousarray(fields, dtype=np.float64)
    if f.size != _lib.lib.azp_param_num_fields(evaluator):
        raise ValueError("wrong number of parameter fields")
    out = np.zeros(param_size(evaluator, bits), dtype=np.uint8)
    _lib.check(_lib.lib.azp_param_pack(evaluator, bits, f.ctypes.data, out.ctypes.data), "param_pack")
    return out


def unpack_params(evaluator, bits, raw):
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    f = np.zeros(_lib.lib.azp_param_num_fields(evaluator), dtype=np.float64)
    _lib.check(_lib.lib.azp_param_unpack(evaluator, bits, raw.ctypes.data, f.ctypes.data), "param_unpack")
    return f


def dpd_alpha(bits, seed, tag_i, tag_j, timestep):
    return float(_lib.lib.azp_dpd_alpha(bits, seed, tag_i, tag_j, timestep))


def philox4x32_10(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    _lib.lib.azp_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def gather_rows(src, idx, out, stream=None):
    """out[k] = src[idx[k]] for 2-D Scalar4 tensors on the GPU (halo packing)."""
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    row_bytes = src.shape[1] * src.element_size()
    rc = _lib.lib.azp_gather_rows(src.data_ptr(), idx.data_ptr(), int(idx.numel()), row_bytes,
                                  out.data_ptr(), ctypes.c_void_p(stream))
    _lib.check(rc, "gather_rows")
    return out


def push_rows(src, idx, dst_addr, stream=None):
    """*(row at dst_addr[k]) = src[idx[k]] -- halo push; ``dst_addr`` (int64 tensor of device
    addresses) may point into peer GPUs' memory (symmetric-memory buffers over NVLink)."""
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    row_bytes = src.shape[1] * src.element_size()
    rc = _lib.lib.azp_push_rows(src.data_ptr(), idx.data_ptr(), dst_addr.data_ptr(),
                                int(idx.numel()), row_bytes, ctypes.c_void_p(stream))
    _lib.check(rc, "push_rows")
