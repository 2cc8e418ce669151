"""Spike-triggered additive plasticity (STDP building blocks) on CSR / CSC / fixed-number / dense
weights -- SURVEY 8(f)-4.

Mirrors ref: _csr/plasticity_binary.py (update_csr_on_binary_pre :45, update_csr_on_binary_post :477,
update_csc_on_binary_pre :968, update_csc_on_binary_post :1066), _fcn/plasticity_binary.py
(update_fixed_post_conn_on_binary_pre :207, update_fixed_pre_conn_on_binary_post :269),
_dense/plasticity_binary.py (update_dense_on_binary_pre :42, update_dense_on_binary_post :360).
Same names, argument order and error behaviour.  Like the reference (functional jax arrays) every
function returns the updated weights as a NEW tensor; `inplace=True` updates the given tensor
instead (no copy of the matrix per step).
"""
import math

import torch

from . import _lib
from ._op import maybe_decimal, split_mantissa_unit
from ._runtime import as_tensor, event_args, ptr, require_cuda, stream_ptr, workspace, wt_code


def _bounds(w_min, w_max):
    lo = -math.inf if w_min is None else float(split_mantissa_unit(w_min)[0])
    hi = math.inf if w_max is None else float(split_mantissa_unit(w_max)[0])
    return lo, hi


def _spike_events(spike):
    """float spikes are active when != 0 here (the operators use > 0): float32 goes to the kernel
    as is (nonzero predicate on the device), other float widths are reduced to bool first."""
    if isinstance(spike, torch.Tensor) and spike.dtype.is_floating_point and spike.dtype != torch.float32:
        spike = spike != 0
    return event_args(spike)


def _check_heterogeneous(weight, indices):
    if weight.numel() == 1 and indices.numel() > 1:
        raise ValueError(
            "Plasticity updates require per-synapse (heterogeneous) weights, but received "
            "a homogeneous (size-1) weight. Materialize per-synapse weights first "
            "(e.g. broadcast to the connectivity shape) before applying a plasticity update.")


def _finish_clip(w, lo, hi):
    """the reference clips the whole array (u.math.clip), not only the touched rows"""
    if lo != -math.inf or hi != math.inf:
        _lib.check(_lib.lib().be_b200_clip(ptr(w), wt_code(w.dtype), w.numel(), lo, hi, stream_ptr(w.device)), "clip")
    return w


def _rows_update(weight, indices, indptr, perm, spike, trace, w_min, w_max, n_rows, n_trace, inplace, name):
    weight, wunit = split_mantissa_unit(weight)
    trace = split_mantissa_unit(trace)[0]
    weight, indices = as_tensor(weight), as_tensor(indices)
    require_cuda(weight, indices, indptr, perm)
    _check_heterogeneous(weight, indices)
    assert weight.dtype.is_floating_point, "Weights must be a floating-point type."
    assert indices.dtype == torch.int32, f"indices must be int32, got {indices.dtype}"
    w = weight if inplace else weight.clone()
    assert w.is_contiguous(), "weights must be contiguous"
    indices = indices.contiguous()
    trace = as_tensor(trace, device=w.device).to(w.dtype).contiguous()
    assert trace.ndim == 1 and trace.shape[0] == n_trace, f"trace length must be {n_trace}, got {tuple(trace.shape)}"
    ev = _spike_events(spike)
    assert ev.batch is None and ev.n == n_rows, f"spike length must be {n_rows}, got {ev.n}"
    lo, hi = _bounds(w_min, w_max)
    n_conn = -1
    if indptr is None:
        assert indices.ndim == 2 and indices.shape[0] == n_rows
        n_conn = int(indices.shape[1])
    else:
        indptr = as_tensor(indptr).contiguous()
        assert indptr.shape[0] == n_rows + 1, f"indptr must have shape ({n_rows + 1},)"
    if perm is not None:
        perm = as_tensor(perm).contiguous()
        assert perm.dtype in (torch.int32, torch.int64)
    ws = workspace(w.device, ev.n, 1)
    rc = _lib.lib().be_b200_plasticity_rows(
        ptr(w), wt_code(w.dtype), ptr(indices), ptr(indptr), int(indptr is not None and indptr.dtype == torch.int64),
        n_conn, ptr(perm), int(perm is not None and perm.dtype == torch.int64), ptr(ev.spikes), ptr(ev.aux), ev.kind,
        ptr(trace), n_rows, n_trace, indices.numel(), lo, hi, ptr(ws), ws.numel(), stream_ptr(w.device))
    _lib.check(rc, name)
    return maybe_decimal(_finish_clip(w, lo, hi), wunit)


def update_csr_on_binary_pre(weight, indices, indptr, pre_spike, post_trace, w_min=None, w_max=None, *, shape,
                             backend=None, inplace=False):
    """W[i, j] <- clip(W[i, j] + post_trace[j]) for every stored synapse of a firing pre neuron i
    (ref: _csr/plasticity_binary.py:45)."""
    return _rows_update(weight, indices, indptr, None, pre_spike, post_trace, w_min, w_max, int(shape[0]),
                        int(shape[1]), inplace, "update_csr_on_binary_pre")


def update_csr_on_binary_post(weight, indices, indptr, weight_indices, pre_trace, post_spike, w_min=None, w_max=None,
                              *, shape, backend=None, inplace=False):
    """W[i, j] <- clip(W[i, j] + pre_trace[i]) for every stored synapse of a firing post neuron j.
    (indices, indptr, weight_indices) is the CSC mirror of the CSR structure (csr_to_csc_index):
    pre ids per post column and the CSR slot of each (ref: _csr/plasticity_binary.py:477)."""
    return _rows_update(weight, indices, indptr, weight_indices, post_spike, pre_trace, w_min, w_max, int(shape[1]),
                        int(shape[0]), inplace, "update_csr_on_binary_post")


def update_csc_on_binary_pre(weight, indices, indptr, pre_spike, post_trace, w_min=None, w_max=None, *, shape,
                             backend=None, inplace=False):
    """CSC-ordered weights, pre-triggered (the unfavourable direction: goes through the CSR mirror of
    the CSC structure; ref: _csr/plasticity_binary.py:968)."""
    from ._csr import csr_to_csc_index
    # (indices, indptr) compress the columns of W [n_pre, n_post] == CSR of W^T [n_post, n_pre]
    csr_indptr, csr_indices, perm = csr_to_csc_index(as_tensor(indptr), as_tensor(indices), shape=tuple(shape[::-1]))
    return update_csr_on_binary_post(weight, csr_indices, csr_indptr, perm, post_trace, pre_spike, w_min, w_max,
                                     shape=tuple(shape[::-1]), backend=backend, inplace=inplace)


def update_csc_on_binary_post(weight, indices, indptr, pre_trace, post_spike, w_min=None, w_max=None, *, shape,
                              backend=None, inplace=False):
    """CSC-ordered weights, post-triggered: the favourable direction (ref: _csr/plasticity_binary.py:1066)."""
    return update_csr_on_binary_pre(weight, indices, indptr, post_spike, pre_trace, w_min, w_max,
                                    shape=tuple(shape[::-1]), backend=backend, inplace=inplace)


def update_fixed_post_conn_on_binary_pre(data, indices, pre_spike, post_trace, w_min=None, w_max=None, *, shape,
                                         backend=None, inplace=False):
    """FixedNumPerPre weights [num_pre, num_conn], pre-triggered (ref: _fcn/plasticity_binary.py:207)."""
    indices = as_tensor(indices)
    assert indices.shape[0] == shape[0], "indices rows must equal num_pre."
    return _rows_update(data, indices, None, None, pre_spike, post_trace, w_min, w_max, int(shape[0]), int(shape[1]),
                        inplace, "update_fixed_post_conn_on_binary_pre")


def update_fixed_pre_conn_on_binary_post(data, indices, pre_trace, post_spike, w_min=None, w_max=None, *, shape,
                                         backend=None, inplace=False):
    """FixedNumPerPost weights [num_post, num_conn] holding pre ids, post-triggered
    (ref: _fcn/plasticity_binary.py:269)."""
    indices = as_tensor(indices)
    assert indices.shape[0] == shape[1], "indices rows must equal num_post."
    return _rows_update(data, indices, None, None, post_spike, pre_trace, w_min, w_max, int(shape[1]), int(shape[0]),
                        inplace, "update_fixed_pre_conn_on_binary_post")


def _dense_update(weight, spike, trace, on_post, w_min, w_max, inplace, name):
    weight, wunit = split_mantissa_unit(weight)
    trace = split_mantissa_unit(trace)[0]
    weight = as_tensor(weight)
    require_cuda(weight)
    assert weight.ndim == 2, f"weight must be 2-D, got {weight.ndim}-D"
    rows, cols = weight.shape
    w = weight if inplace else weight.clone()
    assert w.is_contiguous(), "weights must be contiguous"
    trace = as_tensor(trace, device=w.device).to(w.dtype).contiguous()
    ev = _spike_events(spike)
    n_spk, n_tr = (cols, rows) if on_post else (rows, cols)
    assert ev.batch is None and ev.n == n_spk, f"spike length must be {n_spk}, got {ev.n}"
    assert trace.ndim == 1 and trace.shape[0] == n_tr, f"trace length must be {n_tr}, got {tuple(trace.shape)}"
    lo, hi = _bounds(w_min, w_max)
    ws = workspace(w.device, ev.n, 1)
    rc = _lib.lib().be_b200_plasticity_dense(ptr(w), wt_code(w.dtype), rows, cols, ptr(ev.spikes), ptr(ev.aux), ev.kind,
                                             ptr(trace), int(on_post), lo, hi, ptr(ws), ws.numel(), stream_ptr(w.device))
    _lib.check(rc, name)
    return maybe_decimal(_finish_clip(w, lo, hi), wunit)


def update_dense_on_binary_pre(weight, pre_spike, post_trace, w_min=None, w_max=None, *, backend=None, inplace=False):
    """W[i, :] <- clip(W[i, :] + post_trace) for firing pre neurons i (ref: _dense/plasticity_binary.py:42)."""
    return _dense_update(weight, pre_spike, post_trace, False, w_min, w_max, inplace, "update_dense_on_binary_pre")


def update_dense_on_binary_post(weight, pre_trace, post_spike, w_min=None, w_max=None, *, backend=None, inplace=False):
    """W[:, j] <- clip(W[:, j] + pre_trace) for firing post neurons j (ref: _dense/plasticity_binary.py:360)."""
    return _dense_update(weight, post_spike, pre_trace, True, w_min, w_max, inplace, "update_dense_on_binary_post")
