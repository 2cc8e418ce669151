"""torch.autograd rules of the event matvec operators -- the host mirror's counterpart of the reference's JVP /
transpose rules (ref: _csr/binary.py:656-722 `_csrmv_jvp_weights`, `_csrmv_jvp_v`, `_csrmv_transpose_rule`;
_fcn/binary.py and _dense/binary.py define the same three rules for their primitives).

The operators are linear in the weights, so the cotangent of the weights is itself an event-driven product and
runs on the same CUDA kernels:

* homogeneous weight:  ct_w = <ct_y, op(1)>                       (the reference's scalar branch)
* per-synapse weights: ct_w[slot] = e(pre/col of slot) * ct_y[post/row of slot] -- zero except in the rows
  (columns) that fired, i.e. exactly the spike-triggered row traversal of the plasticity kernels applied to a zero
  array with ct_y as the trace (`update_csr_on_binary_pre/_post`, `update_fixed_post_conn_on_binary_pre`,
  `update_dense_on_binary_pre/_post`);
* float events (the reference treats the product as linear in a float event vector, which is what surrogate-gradient
  training uses): ct_v = op_float(w, ct_y) with the opposite `transpose` (`csrmv`, `fcnmv`, dense matvec).

Bool / bit-packed / compact events carry no gradient.  Only the matvec forms are covered; the matmat forms raise in
backward for per-synapse weights (their weight cotangent is a sampled dense-dense product this library has no kernel
for), homogeneous weights work for both.
"""
import torch

from ._event import BinaryArray, BitPackedBinary, CompactBinary


def _raw(v):
    """dense tensor behind an event container, or None (packed / compact containers have none)"""
    if isinstance(v, BinaryArray):
        return v.value
    if isinstance(v, (BitPackedBinary, CompactBinary)):
        return None
    return v if isinstance(v, torch.Tensor) else None


def _mask(events, ev_tensor):
    """what the plasticity kernels take as the spike argument: they call a float active when != 0, the operators when
    > 0 -- so a float event vector is reduced to the operators' predicate first; containers are passed through"""
    if ev_tensor is None:
        return events
    return (ev_tensor > 0) if ev_tensor.dtype.is_floating_point else ev_tensor


def needs_grad(weights, events):
    ev = _raw(events)
    return torch.is_grad_enabled() and (
        (isinstance(weights, torch.Tensor) and weights.requires_grad) or
        (isinstance(ev, torch.Tensor) and ev.dtype.is_floating_point and ev.requires_grad))


class _CsrMv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, ev_tensor, indices, indptr, events, shape, transpose, backend):
        from ._csr import binary_csrmv_p_call
        ctx.save_for_backward(data.detach(), indices, indptr)
        ctx.events, ctx.shape, ctx.transpose, ctx.backend = events, shape, transpose, backend
        ctx.ev_is_float = ev_tensor is not None and ev_tensor.dtype.is_floating_point
        ctx.mask = _mask(events, None if ev_tensor is None else ev_tensor.detach())
        return binary_csrmv_p_call(data.detach(), indices, indptr, events, shape=shape, transpose=transpose,
                                   backend=backend)[0]

    @staticmethod
    def backward(ctx, g):
        from ._csr import binary_csrmv_p_call, csr_to_csc_index
        from ._float import csrmv
        from ._plasticity import update_csr_on_binary_post, update_csr_on_binary_pre
        data, indices, indptr = ctx.saved_tensors
        g = g.contiguous()
        gw = gv = None
        if ctx.needs_input_grad[0]:
            if data.numel() == 1:      # ref: _csrmv_transpose_rule, scalar branch
                one = torch.ones(1, dtype=data.dtype, device=data.device)
                y1 = binary_csrmv_p_call(one, indices, indptr, ctx.events, shape=ctx.shape, transpose=ctx.transpose,
                                         backend=ctx.backend)[0]
                gw = torch.dot(g.to(y1.dtype), y1).reshape(data.shape)
            elif ctx.transpose:        # ct_w[j] = e(row_j) * g[col_j]: rows that fired, trace = g
                gw = update_csr_on_binary_pre(torch.zeros_like(data), indices, indptr, ctx.mask, g.to(data.dtype),
                                              shape=ctx.shape, inplace=True)
            else:                      # ct_w[j] = e(col_j) * g[row_j]: columns that fired, through the CSC mirror
                csc_indptr, csc_rows, perm = csr_to_csc_index(indptr, indices, shape=ctx.shape)
                gw = update_csr_on_binary_post(torch.zeros_like(data), csc_rows, csc_indptr, perm, g.to(data.dtype),
                                               ctx.mask, shape=ctx.shape, inplace=True)
        if ctx.needs_input_grad[1] and ctx.ev_is_float:     # ref: _csrmv_jvp_v / transpose rule w.r.t. the vector
            gv = csrmv(data, indices, indptr, g.to(data.dtype), shape=ctx.shape, transpose=not ctx.transpose)
        return gw, gv, None, None, None, None, None, None


class _FcnMv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, ev_tensor, indices, events, shape, transpose, backend):
        from ._fcn import binary_fcnmv_p_call
        ctx.save_for_backward(weights.detach(), indices)
        ctx.events, ctx.shape, ctx.transpose, ctx.backend = events, shape, transpose, backend
        ctx.ev_tensor = None if ev_tensor is None else ev_tensor.detach()
        ctx.mask = _mask(events, ctx.ev_tensor)
        return binary_fcnmv_p_call(weights.detach(), indices, events, shape=shape, transpose=transpose,
                                   backend=backend)[0]

    @staticmethod
    def backward(ctx, g):
        from ._fcn import binary_fcnmv_p_call
        from ._float import fcnmv
        from ._plasticity import update_fixed_post_conn_on_binary_pre
        weights, indices = ctx.saved_tensors
        g = g.contiguous()
        gw = gv = None
        ev = ctx.ev_tensor
        if ctx.needs_input_grad[0]:
            if weights.numel() == 1:
                one = torch.ones(1, dtype=weights.dtype, device=weights.device)
                y1 = binary_fcnmv_p_call(one, indices, ctx.events, shape=ctx.shape, transpose=ctx.transpose,
                                         backend=ctx.backend)[0]
                gw = torch.dot(g.to(y1.dtype), y1).reshape(weights.shape)
            elif ctx.transpose:        # scatter: ct_w[i, c] = e(i) * g[indices[i, c]]
                gw = update_fixed_post_conn_on_binary_pre(torch.zeros_like(weights), indices, ctx.mask,
                                                          g.to(weights.dtype), shape=ctx.shape, inplace=True)
            else:                      # gather: ct_w[i, c] = e(indices[i, c]) * g[i] (no row-triggered form: two torch ops)
                if ev is None:
                    raise NotImplementedError("weight gradient of the gather orientation needs a dense event vector")
                act = (ev > 0) if ev.dtype.is_floating_point else (ev != 0)
                gw = act[indices.long()].to(weights.dtype) * g.to(weights.dtype)[:, None]
        if ctx.needs_input_grad[1] and ev is not None and ev.dtype.is_floating_point:
            gv = fcnmv(weights, indices, g.to(weights.dtype), shape=ctx.shape, transpose=not ctx.transpose)
        return gw, gv, None, None, None, None, None


class _DenseMv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, ev_tensor, events, transpose, backend):
        from ._dense import binary_densemv_p_call
        ctx.save_for_backward(weights.detach())
        ctx.events, ctx.transpose = events, transpose
        ctx.ev_tensor = None if ev_tensor is None else ev_tensor.detach()
        ctx.mask = _mask(events, ctx.ev_tensor)
        return binary_densemv_p_call(weights.detach(), events, transpose=transpose, backend=backend)[0]

    @staticmethod
    def backward(ctx, g):
        from ._plasticity import update_dense_on_binary_post, update_dense_on_binary_pre
        (weights,) = ctx.saved_tensors
        g = g.contiguous().to(weights.dtype)
        gw = gv = None
        if ctx.needs_input_grad[0]:
            z = torch.zeros_like(weights)
            # transpose: out[m] = sum_k W[k, m] e(k)  ->  ct_W[k, :] = e(k) g;   else out[k] = sum_m W[k, m] e(m) -> ct_W[:, m] = g e(m)
            gw = (update_dense_on_binary_pre(z, ctx.mask, g, inplace=True) if ctx.transpose
                  else update_dense_on_binary_post(z, g, ctx.mask, inplace=True))
        ev = ctx.ev_tensor
        if ctx.needs_input_grad[1] and ev is not None and ev.dtype.is_floating_point:
            gv = (weights @ g) if ctx.transpose else (weights.t() @ g)     # the dense float matvec is a plain library GEMV
        return gw, gv, None, None, None


def csrmv(data, indices, indptr, events, shape, transpose, backend):
    ev = _raw(events)
    return _CsrMv.apply(data, ev, indices, indptr, events, tuple(shape), bool(transpose), backend)


def fcnmv(weights, indices, events, shape, transpose, backend):
    ev = _raw(events)
    return _FcnMv.apply(weights, ev, indices, events, tuple(shape), bool(transpose), backend)


def densemv(weights, events, transpose, backend):
    ev = _raw(events)
    return _DenseMv.apply(weights, ev, events, bool(transpose), backend)


class _LinearInParams(torch.autograd.Function):
    """y = call(params) is linear in a few scalar parameters (JITC weight parameters: the scalar weight, (w_loc,
    w_scale), (w_low, w_high); ref: the jvp rules of _jit_scalar / _jit_normal / _jit_uniform binary.py): the cotangent
    of parameter i is <ct_y, call(e_i)>, one regenerating product with a unit parameter vector."""

    @staticmethod
    def forward(ctx, call, *params):
        ctx.call = call
        ctx.save_for_backward(*[p.detach() for p in params])
        return call([p.detach() for p in params])

    @staticmethod
    def backward(ctx, g):
        params = ctx.saved_tensors
        grads = []
        for i, p in enumerate(params):
            if not ctx.needs_input_grad[1 + i]:
                grads.append(None)
                continue
            unit = [torch.ones_like(q) if j == i else torch.zeros_like(q) for j, q in enumerate(params)]
            y = ctx.call(unit)
            grads.append((g.to(y.dtype) * y).sum().to(p.dtype).reshape(p.shape))
        return (None, *grads)


def linear_in_params(call, params):
    return _LinearInParams.apply(call, *params)
