"""torch.autograd.Function wrappers over the C ABI (include/xggm_b200.h).

Each Function allocates its outputs / saved buffers with torch (so the caching
allocator and stream semantics stay torch's) and enqueues the kernels on
``torch.cuda.current_stream()`` through ``_lib.call``.  There is no eager
PyTorch fallback anywhere in this file.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import call, f32, ptr, ptr_table

LN_EPS = 1e-5
KIND = {"GCN": 0, "GIN": 1}

# OPT-IN in-place gradient accumulation: when a leaf parameter has been registered with an
# ``xggm_b200.ddp.FlatGrads`` bucket (which tags it ``_xggm_flat``) and its ``.grad`` still is that bucket's
# slice, the backward kernels accumulate into the slice directly and the autograd Function returns None for
# the parameter: no temporary gradient, no extra add kernel.  Parameters that were NOT registered always get
# their gradients through autograd (hooks, torch.autograd.grad, torch DDP's reducer keep working).  For the
# registered ones ``loss.backward()`` is the supported entry; ``torch.autograd.grad`` w.r.t. them returns None
# -- set this flag to False around such calls.
FUSE_GRAD_ACCUMULATION = True


def _grad_target(p):
    """The live .grad slice of ``p`` if the kernels may accumulate into it, else None."""
    if not FUSE_GRAD_ACCUMULATION or not getattr(p, "is_leaf", False) or not p.requires_grad:
        return None
    flat = getattr(p, "_xggm_flat", None)
    if flat is None:
        return None
    g = p.grad
    if g is None or g.dtype != torch.float32 or g.device != p.device or g.shape != p.shape or not g.is_contiguous():
        return None
    lo = flat.flat.data_ptr()
    if not (lo <= g.data_ptr() < lo + flat.flat.numel() * flat.flat.element_size()):
        return None     # un-linked (model.zero_grad()): autograd allocates, FlatGrads.relink() repairs later
    flat.touch(p)
    return g


def _u8(t):
    if t is None:
        return None
    if t.dtype != torch.uint8:
        t = t.to(torch.uint8)
    _lib.check_device(t)
    return t.contiguous()


# ---------------------------------------------------------------------------
# operand-plane hand-over (include/xggm_b200.h, "Operand-plane hand-over")
# ---------------------------------------------------------------------------
# A kernel that produces a [B,N,H] tensor under a tensor-core engine also emits its bf16 operand planes;
# the buffer rides on the tensor object (``_xggm_planes``) and the consumers of that tensor (next GNN layer,
# adjacency regeneration fwd/bwd) pick it up instead of re-splitting the tensor.  The record is dropped as
# soon as the tensor is modified in place, re-allocated, or the engine precision changes.
def _planes_enabled():
    return _lib.load().xggm_get_precision() != _lib.PRECISIONS["fp32_simt"]


def _new_planes(t):
    # (the producers emit planes exactly when the tensor-core engine can address the tensor: H % 8 == 0)
    if not _planes_enabled() or t.shape[-1] % 8 != 0:
        return None
    n = _lib.load().xggm_planes_bytes(t.numel())
    return torch.empty(max(int(n), 16), device=t.device, dtype=torch.uint8)


def _attach_planes(t, planes):
    if planes is not None:
        t._xggm_planes = (planes, t._version, t.data_ptr(), _lib.load().xggm_get_precision())


def _planes_of(t):
    rec = getattr(t, "_xggm_planes", None)
    if rec is None or not t.is_contiguous() or t.dtype != torch.float32:
        return None
    planes, version, data_ptr, prec = rec
    if t._version != version or t.data_ptr() != data_ptr or prec != _lib.load().xggm_get_precision():
        return None
    return planes


# ---------------------------------------------------------------------------
# prepared weight planes (include/xggm_b200.h, "Prepared weight planes")
# ---------------------------------------------------------------------------
# A weight matrix changes once per optimiser step but is read as a tensor-core operand by several products per
# step (W in the forward pass, W^T in the input-gradient product, in every layer call).  For parameters that
# OPT IN -- ``cache_weight_planes(params)``; ``xggm_b200.optim.BertAdam`` opts its parameters in and rebuilds the
# planes right after its update kernel -- both bf16 layouts are built once (one launch for all stale matrices)
# and handed to the entry points, instead of being re-split inside every call.  Opt-in because a record cannot
# see writes that bypass torch's version counter (``p.data.add_()``, as the reference's own optimiser does): code
# that updates weights that way must call ``refresh_weight_planes`` itself or leave the cache off.
def cache_weight_planes(params, enable=True):
    for p in params:
        if enable:
            p._xggm_wp_ok = True
        else:
            p._xggm_wp_ok = False
            p._xggm_wp = None


def _wp_valid(w):
    rec = getattr(w, "_xggm_wp", None)
    if rec is None:
        return None
    buf, version, data_ptr, prec = rec
    if w._version != version or w.data_ptr() != data_ptr or prec != _lib.load().xggm_get_precision():
        return None
    return buf


def _build_weight_planes(ws):
    lib = _lib.load()
    prec = lib.xggm_get_precision()
    n = len(ws)
    bufs = []
    for w in ws:
        N, K = w.shape
        nbytes = int(lib.xggm_weight_planes_bytes(N, K))
        rec = getattr(w, "_xggm_wp", None)
        buf = rec[0] if (rec is not None and rec[0].numel() == nbytes and rec[0].device == w.device) else \
            torch.empty(nbytes, device=w.device, dtype=torch.uint8)
        bufs.append(buf)
    wt, bt = ptr_table(ws), ptr_table(bufs)
    Ns = (C.c_int * n)(*[int(w.shape[0]) for w in ws])
    Ks = (C.c_int * n)(*[int(w.shape[1]) for w in ws])
    call("xggm_weight_planes_build", wt, bt, Ns, Ks, n)
    for w, buf in zip(ws, bufs):
        w._xggm_wp = (buf, w._version, w.data_ptr(), prec)


def weight_planes(ws):
    """Prepared plane buffers of the 2-D fp32 weights ``ws`` (None where a weight has not opted in or the engine
    cannot use them); stale ones are rebuilt first, all in one launch."""
    if not _planes_enabled():
        return [None] * len(ws)
    ok = [bool(getattr(w, "_xggm_wp_ok", False)) and w.dim() == 2 and w.shape[1] % 8 == 0 and w.is_contiguous()
          and w.dtype == torch.float32 and w.is_cuda for w in ws]
    stale = [w for w, o in zip(ws, ok) if o and _wp_valid(w) is None]
    if stale:
        _build_weight_planes(stale)
    return [w._xggm_wp[0] if o else None for w, o in zip(ws, ok)]


def refresh_weight_planes(params):
    """Rebuild (in place, one launch) the planes of every parameter in ``params`` that has a plane record: to be
    called after the weights were updated by something torch's version counter does not see."""
    if not _planes_enabled():
        return
    ws = [p for p in params if getattr(p, "_xggm_wp_ok", False) and getattr(p, "_xggm_wp", None) is not None]
    if ws:
        _build_weight_planes(ws)


# ---------------------------------------------------------------------------
# dense projection (nn.Linear)
# ---------------------------------------------------------------------------
def _linear_work(M, N, K, device):
    """Scratch for the bf16 operand planes of the tcgen05 engine (caller-owned, see xggm_b200.h)."""
    n = _lib.load().xggm_linear_work_bytes(M, N, K)
    return torch.empty(max(int(n), 16), device=device, dtype=torch.uint8)


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, w, bias, resid):
        a2 = f32(a, "input").reshape(-1, a.shape[-1])
        w = f32(w, "weight")
        M, K = a2.shape
        N = w.shape[0]
        if w.shape[1] != K:
            raise RuntimeError(f"xggm_b200.linear: weight {tuple(w.shape)} does not match input width {K}")
        bias = None if bias is None else f32(bias, "bias")
        r2 = None if resid is None else f32(resid, "resid").reshape(-1, N)
        out = torch.empty((M, N), device=a.device, dtype=torch.float32)
        work = _linear_work(M, N, K, a.device)
        wp = weight_planes([w])[0] if M > 0 else None
        call("xggm_linear_fwd_ex", ptr(a2), ptr(w), ptr(bias), ptr(r2), ptr(out), M, N, K, ptr(work), ptr(wp))
        ctx.w_planes = wp
        ctx.save_for_backward(a2, w)
        ctx.bias_ref = bias   # only consulted for its .grad buffer in backward
        ctx.has_bias, ctx.has_resid, ctx.in_shape = bias is not None, resid is not None, a.shape
        return out.reshape(*a.shape[:-1], N)

    @staticmethod
    def backward(ctx, g):
        a2, w = ctx.saved_tensors
        M, K = a2.shape
        N = w.shape[0]
        g2 = f32(g).reshape(M, N)
        ga = gw = gb = gr = None
        work = _linear_work(M, N, K, g.device)
        if ctx.needs_input_grad[0]:
            ga = torch.empty_like(a2)
            wp = ctx.w_planes if (ctx.w_planes is not None and _wp_valid(w) is ctx.w_planes) else None
            call("xggm_linear_bwd_input_ex", ptr(g2), ptr(w), ptr(ga), M, N, K, 0, ptr(work), ptr(wp))
            ga = ga.reshape(ctx.in_shape)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            bias = ctx.bias_ref
            tw, tb = _grad_target(w), (_grad_target(bias) if ctx.has_bias else None)
            if tw is not None and (not ctx.has_bias or tb is not None):
                call("xggm_linear_bwd_weight", ptr(g2), ptr(a2), ptr(tw), ptr(tb), M, N, K, 1, ptr(work))
            else:
                gw = torch.empty_like(w)
                gb = torch.empty(N, device=w.device, dtype=torch.float32) if ctx.has_bias else None
                call("xggm_linear_bwd_weight", ptr(g2), ptr(a2), ptr(gw), ptr(gb), M, N, K, 0, ptr(work))
        if ctx.has_resid and ctx.needs_input_grad[3]:
            gr = g
        return ga, gw, gb, gr


def linear(a, w, bias=None, resid=None):
    """a w^T + bias + resid through the library's GEMM engine."""
    return _Linear.apply(a, w, bias, resid)


# ---------------------------------------------------------------------------
# message passing
# ---------------------------------------------------------------------------
def _adj_work(B, N, H, device):
    """Scratch for the tensor-core message-passing path (node planes + block-diagonal coefficient tiles)."""
    n = _lib.load().xggm_adj_apply_work_bytes(B, N, H)
    return torch.empty(max(int(n), 16), device=device, dtype=torch.uint8)


class _AdjApply(torch.autograd.Function):
    @staticmethod
    def forward(ctx, adj, x, alpha0, alpha_dev, self_w):
        adj, x = f32(adj, "adj"), f32(x, "x")
        B, N, H = x.shape
        if adj.shape != (B, N, N):
            raise RuntimeError(f"xggm_b200.adj_apply: adj {tuple(adj.shape)} vs x {tuple(x.shape)}")
        al = None if alpha_dev is None else f32(alpha_dev, "alpha")
        out = torch.empty_like(x)
        tc_work = _adj_work(B, N, H, x.device)    # (scratch buffers are named locals: they must outlive the call)
        call("xggm_adj_apply_fwd", ptr(adj), ptr(x), ptr(out), B, N, H, float(alpha0), ptr(al), float(self_w),
             ptr(tc_work))
        ctx.save_for_backward(adj, x, al)
        ctx.cfg = (float(alpha0), float(self_w))
        return out

    @staticmethod
    def backward(ctx, g):
        adj, x, al = ctx.saved_tensors
        alpha0, self_w = ctx.cfg
        B, N, H = x.shape
        g = f32(g)
        gx = torch.empty_like(x)
        graw = torch.empty_like(adj)
        tc_work = _adj_work(B, N, H, x.device)
        call("xggm_adj_apply_bwd", ptr(adj), ptr(x), ptr(g), ptr(gx), ptr(graw), B, N, H, alpha0, ptr(al),
             self_w, 0, ptr(tc_work))
        alpha = alpha0 if al is None else alpha0 + al
        gadj = graw * alpha
        gal = None
        if al is not None and ctx.needs_input_grad[3]:
            gal = (graw * adj).sum().reshape(al.shape)
        return gadj, gx, None, gal, None


def adj_apply(adj, x, alpha0=1.0, alpha_dev=None, self_w=0.0):
    """self_w*x + (alpha0 + alpha_dev) * adj @ x."""
    return _AdjApply.apply(adj, x, alpha0, alpha_dev, self_w)


# ---------------------------------------------------------------------------
# row ops
# ---------------------------------------------------------------------------
class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, gamma, beta, eps):
        u2 = f32(u, "input").reshape(-1, u.shape[-1])
        gamma, beta = f32(gamma, "weight"), f32(beta, "bias")
        M, H = u2.shape
        h = torch.empty_like(u2)
        xhat = torch.empty_like(u2)
        rstd = torch.empty(M, device=u.device, dtype=torch.float32)
        call("xggm_layernorm_fwd", ptr(u2), ptr(gamma), ptr(beta), ptr(h), ptr(xhat), ptr(rstd), M, H, float(eps))
        ctx.save_for_backward(xhat, rstd, gamma)
        ctx.beta_ref = beta
        return h.reshape(u.shape)

    @staticmethod
    def backward(ctx, g):
        xhat, rstd, gamma = ctx.saved_tensors
        M, H = xhat.shape
        g2 = f32(g).reshape(M, H)
        gu = torch.empty_like(xhat)
        tg, tb = _grad_target(gamma), _grad_target(ctx.beta_ref)
        fused = tg is not None and tb is not None
        gg = tg if fused else torch.zeros(H, device=g.device, dtype=torch.float32)
        gb = tb if fused else torch.zeros(H, device=g.device, dtype=torch.float32)
        call("xggm_layernorm_bwd", ptr(g2), ptr(xhat), ptr(rstd), ptr(gamma), ptr(gu), ptr(gg), ptr(gb), M, H)
        return gu.reshape(g.shape), (None if fused else gg), (None if fused else gb), None


def layer_norm(u, gamma, beta, eps=LN_EPS):
    return _LayerNorm.apply(u, gamma, beta, eps)


class _GeluLnDrop(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, gamma, beta, keep, scale, eps):
        z2 = f32(z, "input").reshape(-1, z.shape[-1])
        gamma, beta = f32(gamma, "weight"), f32(beta, "bias")
        keep = _u8(keep)
        M, H = z2.shape
        out = torch.empty_like(z2)
        mean = torch.empty(M, device=z.device, dtype=torch.float32)
        rstd = torch.empty(M, device=z.device, dtype=torch.float32)
        call("xggm_gelu_ln_drop_fwd", ptr(z2), ptr(gamma), ptr(beta), ptr(keep), float(scale), ptr(out),
             ptr(mean), ptr(rstd), M, H, float(eps), 0)
        ctx.save_for_backward(z2, mean, rstd, gamma, keep)
        ctx.beta_ref = beta
        ctx.scale = float(scale)
        return out.reshape(z.shape)

    @staticmethod
    def backward(ctx, g):
        z2, mean, rstd, gamma, keep = ctx.saved_tensors
        M, H = z2.shape
        g2 = f32(g).reshape(M, H)
        gz = torch.empty_like(z2)
        tg, tb = _grad_target(gamma), _grad_target(ctx.beta_ref)
        fused = tg is not None and tb is not None
        gg = tg if fused else torch.zeros(H, device=g.device, dtype=torch.float32)
        gb = tb if fused else torch.zeros(H, device=g.device, dtype=torch.float32)
        call("xggm_gelu_ln_drop_bwd", ptr(g2), ptr(z2), ptr(mean), ptr(rstd), ptr(gamma), ptr(keep), ctx.scale,
             ptr(gz), ptr(gg), ptr(gb), M, H)
        return gz.reshape(g.shape), (None if fused else gg), (None if fused else gb), None, None, None


def gelu_ln_drop(z, gamma, beta, keep=None, drop_p=0.0, eps=LN_EPS):
    """dropout(LayerNorm(GeLU_erf(z))) with an explicit keep-mask (None = no dropout)."""
    scale = 1.0 / (1.0 - drop_p) if keep is not None else 1.0
    return _GeluLnDrop.apply(z, gamma, beta, keep, scale, eps)


# ---------------------------------------------------------------------------
# adjacency regeneration
# ---------------------------------------------------------------------------
class _AdjRegen(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, squash, x_planes):
        x = f32(x, "x")
        B, N, H = x.shape
        adj = torch.empty((B, N, N), device=x.device, dtype=torch.float32)
        S = torch.empty_like(adj)
        amax = torch.empty((B, N), device=x.device, dtype=torch.int32)
        work = None
        if x_planes is None:
            nwork = _lib.load().xggm_adj_regen_work_bytes(B, N, H)
            work = torch.empty(max(int(nwork), 16), device=x.device, dtype=torch.uint8)
        call("xggm_adj_regen_fwd_ex", ptr(x), ptr(adj), ptr(S), ptr(amax), B, N, H, int(squash), ptr(work), ptr(x_planes))
        ctx.save_for_backward(x, S, amax)
        ctx.x_planes = x_planes
        ctx.squash = int(squash)
        ctx.mark_non_differentiable(amax)
        ctx.set_materialize_grads(False)   # no zero-filled stand-in for the (never used) gradient of amax
        return adj, amax

    @staticmethod
    def backward(ctx, g, _g_amax):
        x, S, amax = ctx.saved_tensors
        B, N, H = x.shape
        if g is None:
            return None, None, None
        g = f32(g)
        gx = torch.empty_like(x)
        work, tc_work = torch.empty_like(S), _adj_work(B, N, H, x.device)
        call("xggm_adj_regen_bwd_ex", ptr(g), ptr(x), ptr(S), ptr(amax), ptr(gx), ptr(work), B, N, H, ctx.squash, 0,
             ptr(tc_work), ptr(ctx.x_planes))
        return gx, None, None


def adj_regen(x, squash=True, return_argmax=False):
    """sigmoid(x x^T / colmax) with zero diagonal (ggm.py:225-228)."""
    adj, amax = _AdjRegen.apply(x, squash, _planes_of(x))
    return (adj, amax) if return_argmax else adj


# ---------------------------------------------------------------------------
# whole GCN / GIN layer
# ---------------------------------------------------------------------------
class _GnnLayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind, n_convs, drop_p, keeps, philox, x_planes, out_planes, x, adj, *params):
        # adj: the [B,N,N] adjacency, or ("regen", squash): the adjacency is REGENERATED from x inside this node
        # (ggm.py:225-228 followed by the next layer, ggm.py:221-222) -- x then has one consumer instead of two, so
        # autograd needs no add kernel for its gradient: the regeneration's backward accumulates into the layer's gx
        x = f32(x, "x")
        B, N, H = x.shape
        regen = None
        if isinstance(adj, tuple):
            regen = int(bool(adj[1]))
            adj = torch.empty((B, N, N), device=x.device, dtype=torch.float32)
            S = torch.empty_like(adj)
            amax = torch.empty((B, N), device=x.device, dtype=torch.int32)
            rwork = None
            if x_planes is None:
                rwork = torch.empty(max(int(_lib.load().xggm_adj_regen_work_bytes(B, N, H)), 16), device=x.device, dtype=torch.uint8)
            call("xggm_adj_regen_fwd_ex", ptr(x), ptr(adj), ptr(S), ptr(amax), B, N, H, regen, ptr(rwork), ptr(x_planes))
            ctx.regen_saved = (S, amax)
        ctx.regen = regen
        adj = f32(adj, "adj")
        if adj.shape != (B, N, N):
            raise RuntimeError(f"xggm_b200: adj {tuple(adj.shape)} does not match x {tuple(x.shape)}")
        per_conv = 3 if kind == 0 else 5
        n_cp = per_conv * n_convs
        params = [f32(p, "parameter") for p in params]
        if len(params) != n_cp + 4 * (n_convs + 1):
            raise RuntimeError("xggm_b200: wrong number of layer parameters")
        cp, hp = params[:n_cp], params[n_cp:]
        keeps = None if keeps is None else [_u8(k) for k in keeps]
        if keeps is not None and (len(keeps) != n_convs + 1 or any(k.numel() != B * N * H for k in keeps)):
            raise RuntimeError("xggm_b200: need one [B,N,H] keep-mask per read-out head")
        lib = _lib.load()
        n_saved = lib.xggm_gnn_saved_floats(kind, B, N, H, n_convs)
        n_work = lib.xggm_gnn_work_floats(kind, B, N, H, n_convs)
        saved = torch.empty(max(n_saved, 1), device=x.device, dtype=torch.float32)
        work = torch.empty(max(n_work, 1), device=x.device, dtype=torch.float32)
        out = torch.empty_like(x)
        cpt, hpt = ptr_table(cp), ptr_table(hp)
        kt = None if keeps is None else ptr_table(keeps)
        mats = _layer_matrices(kind, n_convs, cp, hp)
        wps = weight_planes(mats) if B * N > 0 else [None]
        wps = wps if all(b is not None for b in wps) else None
        call("xggm_gnn_fwd_ex", kind, ptr(x), ptr(adj), cpt, hpt, kt, _philox_arg(philox), float(drop_p), ptr(out),
             ptr(saved), ptr(work), ptr(x_planes), ptr(out_planes), B, N, H, n_convs,
             None if wps is None else ptr_table(wps))
        ctx.wps = wps
        ctx.save_for_backward(x, adj, saved, *params)
        ctx.x_planes = x_planes
        ctx.keeps = keeps
        ctx.philox = philox
        ctx.cfg = (kind, n_convs, float(drop_p), n_cp)
        return out

    @staticmethod
    def backward(ctx, g):
        x, adj, saved, *params = ctx.saved_tensors
        kind, n_convs, drop_p, n_cp = ctx.cfg
        B, N, H = x.shape
        cp, hp = params[:n_cp], params[n_cp:]
        g = f32(g)
        lib = _lib.load()
        work = torch.empty(max(lib.xggm_gnn_work_floats(kind, B, N, H, n_convs), 1), device=x.device,
                           dtype=torch.float32)
        gx = torch.empty_like(x)
        # ctx.needs_input_grad: (kind, n_convs, drop_p, keeps, philox, x_planes, out_planes, x, adj, *params);
        # GIN needs gq h^T for d eps
        need_gadj = ctx.needs_input_grad[8] or kind != 0 or ctx.regen is not None
        gadj = torch.empty_like(adj) if need_gadj else None
        targets = [_grad_target(p) for p in params]
        fused = all(t is not None for t in targets)
        grads = targets if fused else [torch.empty_like(p) for p in params]
        kt = None if ctx.keeps is None else ptr_table(ctx.keeps)
        wps = ctx.wps     # the forward pass's prepared weight planes, if the weights have not changed since
        if wps is not None:
            mats = _layer_matrices(kind, n_convs, cp, hp)
            if not all(_wp_valid(w) is b for w, b in zip(mats, wps)):
                wps = None
        call("xggm_gnn_bwd_ex", kind, ptr(g), ptr(x), ptr(adj), ptr_table(cp), ptr_table(hp), kt,
             _philox_arg(ctx.philox), drop_p, ptr(saved), ptr(work), ptr(gx), ptr(gadj), ptr_table(grads[:n_cp]),
             ptr_table(grads[n_cp:]), int(fused), ptr(ctx.x_planes), B, N, H, n_convs,
             None if wps is None else ptr_table(wps))
        if fused:
            grads = [None] * len(params)
        if ctx.regen is not None:   # gx += (dS + dS^T) x : the regenerated adjacency's path back to x
            S, amax = ctx.regen_saved
            d_scratch, tc_work = torch.empty_like(S), _adj_work(B, N, H, x.device)   # (named: they must outlive the call)
            call("xggm_adj_regen_bwd_ex", ptr(gadj), ptr(x), ptr(S), ptr(amax), ptr(gx), ptr(d_scratch), B, N, H,
                 ctx.regen, 1, ptr(tc_work), ptr(ctx.x_planes))
            return (None, None, None, None, None, None, None, gx, None, *grads)
        return (None, None, None, None, None, None, None, gx, (gadj if ctx.needs_input_grad[8] else None), *grads)


def _layer_matrices(kind, n_convs, cp, hp):
    """The 2*n_convs + 1 weight matrices of a layer in the order of the C ABI's weight_planes table."""
    per = 3 if kind == 0 else 5
    return [cp[per * k + (0 if kind == 0 else 1)] for k in range(n_convs)] + [hp[4 * j] for j in range(n_convs + 1)]


def _philox_arg(philox):
    """(seed, stream0, epoch_tensor|None) -> pointer to an xggm_philox_t (or NULL)."""
    if philox is None:
        return None
    seed, stream0, epoch = philox
    spec = _lib.PhiloxSpec(seed & (2 ** 64 - 1), stream0, None if epoch is None else epoch.data_ptr())
    return C.cast(C.pointer(spec), C.c_void_p)


def gnn_layer(kind, x, adj, conv_params, head_params, training=False, drop_p=0.5, regen=None):
    """One GCN (kind='GCN') or GIN (kind='GIN') layer: conv chain + jump-knowledge heads.
    conv_params / head_params are flat lists in the order documented in xggm_b200.h.
    Training-mode dropout of the heads: injected keep-masks if any are queued (parity runs),
    otherwise Philox bits drawn inside the kernels (no mask tensors)."""
    k = KIND[kind]
    per_conv = 3 if k == 0 else 5
    n_convs = len(conv_params) // per_conv
    keeps = philox = None
    if training and drop_p > 0.0:
        if _mask_feed:
            keeps = [keep_mask(x.shape, drop_p, x.device) for _ in range(n_convs + 1)]
        else:
            philox = (torch.initial_seed(), _next_sites(n_convs + 1), _drop.epoch)
    out_planes = _new_planes(x)
    if regen is not None:      # adjacency regenerated from x inside the node (see _GnnLayer.forward)
        adj = ("regen", bool(regen))
    out = _GnnLayer.apply(k, n_convs, drop_p, keeps, philox, _planes_of(x), out_planes, x, adj, *conv_params, *head_params)
    _attach_planes(out, out_planes)
    return out


# ---------------------------------------------------------------------------
# GAT attention
# ---------------------------------------------------------------------------
class _GatAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, a, adj, slope, apply_elu):
        h, adj = f32(h, "h"), f32(adj, "adj")
        ctx.a_shape = a.shape
        a = f32(a, "attn weight").reshape(-1)
        B, N, H = h.shape
        if a.numel() != 2 * H:
            raise RuntimeError("xggm_b200.gat_attn: attention vector must have 2*H entries")
        out = torch.empty_like(h)
        pre = torch.empty_like(h)
        att = torch.empty((B, N, N), device=h.device, dtype=torch.float32)
        call("xggm_gat_attn_fwd", ptr(h), ptr(a), ptr(adj), ptr(out), ptr(att), ptr(pre), B, N, H, float(slope),
             int(apply_elu))
        ctx.save_for_backward(h, a, adj, att, pre)
        ctx.slope, ctx.apply_elu = float(slope), int(apply_elu)
        return out

    @staticmethod
    def backward(ctx, g):
        h, a, adj, att, pre = ctx.saved_tensors
        B, N, H = h.shape
        g = f32(g)
        gh = torch.empty_like(h)
        ga = torch.zeros_like(a)
        work = torch.empty(B * N * H + B * N * N, device=h.device, dtype=torch.float32)
        call("xggm_gat_attn_bwd", ptr(g), ptr(h), ptr(a), ptr(adj), ptr(att), ptr(pre), ptr(gh), ptr(ga),
             ptr(work), B, N, H, ctx.slope, ctx.apply_elu)
        return gh, ga.reshape(ctx.a_shape), None, None, None


def gat_attn(h, a, adj, slope=0.2, apply_elu=True):
    """Masked dense attention of GATConv on already-projected features h (gat.py:31-49)."""
    return _GatAttn.apply(h, a, adj, slope, apply_elu)


class _Gelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = f32(x, "x")
        y = torch.empty_like(x)
        call("xggm_gelu_fwd", ptr(x), ptr(y), x.numel())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = f32(g)
        gx = torch.empty_like(x)
        call("xggm_gelu_bwd", ptr(g), ptr(x), ptr(gx), x.numel())
        return gx


def gelu(x):
    """Exact-erf GeLU (src/lxrt/modeling.py:116-124)."""
    return _Gelu.apply(x)


class _MaskScale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, keep, scale):
        x, keep = f32(x, "x"), _u8(keep)
        y = torch.empty_like(x)
        call("xggm_mask_scale", ptr(x), ptr(keep), float(scale), ptr(y), x.numel())
        ctx.save_for_backward(keep)
        ctx.scale = float(scale)
        return y

    @staticmethod
    def backward(ctx, g):
        (keep,) = ctx.saved_tensors
        g = f32(g)
        gx = torch.empty_like(g)
        call("xggm_mask_scale", ptr(g), ptr(keep), ctx.scale, ptr(gx), g.numel())
        return gx, None, None


class _Avg2Drop(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, keep, scale):
        x, y = f32(x, "x"), f32(y, "y")
        if x.shape != y.shape:
            raise RuntimeError("xggm_b200.avg2_dropout: shape mismatch")
        keep = _u8(keep)
        out = torch.empty_like(x)
        call("xggm_avg2_drop", ptr(x), ptr(y), ptr(keep), float(scale), ptr(out), x.numel())
        ctx.save_for_backward(keep)
        ctx.scale = float(scale)
        return out

    @staticmethod
    def backward(ctx, g):
        (keep,) = ctx.saved_tensors
        g = f32(g)
        gx = torch.empty_like(g)
        call("xggm_mask_scale", ptr(g), ptr(keep), 0.5 * ctx.scale, ptr(gx), g.numel())
        return gx, gx, None, None


class _VisnTail(torch.autograd.Function):
    """dropout((LN(z) + LN(box_fc(boxes))) / 2) in one pass (xggm_visn_tail_*): the tail of VisualFeatEncoder."""

    @staticmethod
    def forward(ctx, z, boxes, bw, bb, g1, b1, g2, b2, keep, scale, eps):
        z2 = f32(z, "z").reshape(-1, z.shape[-1])
        bx = f32(boxes, "boxes").reshape(-1, boxes.shape[-1])
        bw, bb, g1, b1, g2, b2 = (f32(t, "parameter") for t in (bw, bb, g1, b1, g2, b2))
        keep = _u8(keep)
        M, H = z2.shape
        out, xhat1 = torch.empty_like(z2), torch.empty_like(z2)
        rstd1, mean2, rstd2 = (torch.empty(M, device=z.device, dtype=torch.float32) for _ in range(3))
        call("xggm_visn_tail_fwd", ptr(z2), ptr(bx), ptr(bw), ptr(bb), ptr(g1), ptr(b1), ptr(g2), ptr(b2), ptr(keep), float(scale),
             ptr(out), ptr(xhat1), ptr(rstd1), ptr(mean2), ptr(rstd2), M, H, float(eps))
        ctx.save_for_backward(xhat1, rstd1, mean2, rstd2, bx, bw, bb, g1, g2, keep)
        ctx.refs = (b1, b2)
        ctx.scale, ctx.shapes = float(scale), (z.shape, boxes.shape)
        return out.reshape(z.shape)

    @staticmethod
    def backward(ctx, g):
        xhat1, rstd1, mean2, rstd2, bx, bw, bb, g1, g2, keep = ctx.saved_tensors
        b1, b2 = ctx.refs
        M, H = xhat1.shape
        g2d = f32(g).reshape(M, H)
        gz, gt = torch.empty_like(xhat1), torch.empty_like(xhat1)
        tg = [_grad_target(t) for t in (g1, b1, g2, b2)]
        fused = all(t is not None for t in tg)
        gg1, gb1, gg2, gb2 = tg if fused else [torch.zeros(H, device=g.device, dtype=torch.float32) for _ in range(4)]
        call("xggm_visn_tail_bwd", ptr(g2d), ptr(xhat1), ptr(rstd1), ptr(bx), ptr(bw), ptr(bb), ptr(mean2), ptr(rstd2), ptr(g1),
             ptr(g2), ptr(keep), ctx.scale, ptr(gz), ptr(gt), ptr(gg1), ptr(gb1), ptr(gg2), ptr(gb2), M, H)
        # box_fc's own gradients: gW = gt^T boxes [H,4], gb = column sums of gt (K = 4: the exact SIMT product)
        tw, tb = _grad_target(bw), _grad_target(bb)
        if tw is not None and tb is not None:
            call("xggm_linear_bwd_weight", ptr(gt), ptr(bx), ptr(tw), ptr(tb), M, H, bx.shape[1], 1, None)
            gbw = gbb = None
        else:
            gbw, gbb = torch.empty_like(bw), torch.empty_like(bb)
            call("xggm_linear_bwd_weight", ptr(gt), ptr(bx), ptr(gbw), ptr(gbb), M, H, bx.shape[1], 0, None)
        gboxes = None
        if ctx.needs_input_grad[1]:
            gboxes = torch.empty_like(bx)
            call("xggm_linear_bwd_input", ptr(gt), ptr(bw), ptr(gboxes), M, H, bx.shape[1], 0, None)
            gboxes = gboxes.reshape(ctx.shapes[1])
        none4 = (None, None, None, None)
        return (gz.reshape(ctx.shapes[0]), gboxes, gbw, gbb) + (none4 if fused else (gg1, gb1, gg2, gb2)) + (None, None, None)


def visn_tail_supported(H, pos_dim):
    return bool(_lib.load().xggm_visn_tail_supported(int(H), int(pos_dim)))


def visn_tail(z, boxes, box_w, box_b, g1, b1, g2, b2, p, training, eps):
    """dropout((LN(z) g1 + b1 + LN(boxes box_w^T + box_b) g2 + b2) / 2) -- src/lxrt/modeling.py:546-556."""
    if training and p > 0.0:
        return _VisnTail.apply(z, boxes, box_w, box_b, g1, b1, g2, b2, keep_mask(z.shape, p, z.device), 1.0 / (1.0 - p), eps)
    return _VisnTail.apply(z, boxes, box_w, box_b, g1, b1, g2, b2, None, 1.0, eps)


def avg2_dropout(x, y, p, training):
    """dropout((x + y) / 2) -- the tail of VisualFeatEncoder (src/lxrt/modeling.py:553-555)."""
    if training and p > 0.0:
        return _Avg2Drop.apply(x, y, keep_mask(x.shape, p, x.device), 1.0 / (1.0 - p))
    return _Avg2Drop.apply(x, y, None, 1.0)


def dropout(x, p, training):
    """F.dropout with a library-generated (or injected) keep-mask."""
    if not training or p == 0.0:
        return x
    keep = keep_mask(x.shape, p, x.device)
    return _MaskScale.apply(x, keep, 1.0 / (1.0 - p))


# ---------------------------------------------------------------------------
# trainer glue
# ---------------------------------------------------------------------------
class _StripDiag(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a):
        a = f32(a, "adj")
        B, N, _ = a.shape
        out = torch.empty_like(a)
        call("xggm_strip_diag", ptr(a), ptr(out), B, N)
        return out

    @staticmethod
    def backward(ctx, g):
        g = f32(g)
        out = torch.empty_like(g)
        call("xggm_strip_diag", ptr(g), ptr(out), g.shape[0], g.shape[1])
        return out


def strip_diag(a):
    return _StripDiag.apply(a)


class _TriuScatter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, n):
        v = f32(v, "v")
        B = v.shape[0]
        if v.shape[1] != n * (n - 1) // 2:
            raise RuntimeError(f"xggm_b200.triu_scatter: need {n * (n - 1) // 2} values per graph, got {v.shape[1]}")
        adj = torch.empty((B, n, n), device=v.device, dtype=torch.float32)
        call("xggm_triu_scatter_fwd", ptr(v), ptr(adj), B, n)
        ctx.n = n
        return adj

    @staticmethod
    def backward(ctx, g):
        g = f32(g)
        B, n = g.shape[0], ctx.n
        gv = torch.empty((B, n * (n - 1) // 2), device=g.device, dtype=torch.float32)
        call("xggm_triu_scatter_bwd", ptr(g), ptr(gv), B, n)
        return gv, None


def triu_scatter(v, n):
    return _TriuScatter.apply(v, n)


class _EdgeNoise(torch.autograd.Function):
    @staticmethod
    def forward(ctx, adj, randn, sigma):
        adj, randn = f32(adj, "adj"), f32(randn, "randn")
        B, N, _ = adj.shape
        noisy, target = torch.empty_like(adj), torch.empty_like(adj)
        call("xggm_edge_noise", ptr(adj), ptr(randn), float(sigma), ptr(noisy), ptr(target), B, N)
        ctx.mark_non_differentiable(target)
        ctx.set_materialize_grads(False)   # autograd would otherwise fill a [B,N,N] zero gradient for `target`
        return noisy, target

    @staticmethod
    def backward(ctx, g, _gt):
        return g, None, None


class _FeatNoise(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, randn, sigma, planes=None):
        f, randn = f32(f, "feats"), f32(randn, "randn")
        B, N, H = randn.shape
        bcast = f.dim() == 2
        noisy, target = torch.empty_like(randn), torch.empty_like(randn)
        call("xggm_feat_noise_ex", ptr(f), ptr(randn), float(sigma), ptr(noisy), ptr(target), ptr(planes), B, N, H,
             int(bcast))
        ctx.bcast = bcast
        ctx.mark_non_differentiable(target)
        ctx.set_materialize_grads(False)   # autograd would otherwise fill a [B,N,H] zero gradient for `target`
        return noisy, target

    @staticmethod
    def backward(ctx, g, _gt):
        if g is None or not ctx.bcast:
            return g, None, None, None
        g = f32(g)
        B, N, H = g.shape
        out = torch.empty((B, H), device=g.device, dtype=torch.float32)
        call("xggm_sum_nodes", ptr(g), ptr(out), B, N, H)
        return out, None, None, None


class _FeatNoisePhilox(torch.autograd.Function):
    """add_feature_noise_v2 with the Gaussian draw inside the kernel (xggm_feat_noise_philox): no randn tensor."""

    @staticmethod
    def forward(ctx, f, shape, sigma, planes, philox):
        f = f32(f, "feats")
        B, N, H = shape
        bcast = f.dim() == 2
        noisy = torch.empty(shape, device=f.device, dtype=torch.float32)
        target = torch.empty_like(noisy)
        call("xggm_feat_noise_philox", ptr(f), _philox_arg(philox), float(sigma), ptr(noisy), ptr(target), ptr(planes),
             B, N, H, int(bcast))
        ctx.bcast = bcast
        ctx.mark_non_differentiable(target)
        ctx.set_materialize_grads(False)
        return noisy, target

    @staticmethod
    def backward(ctx, g, _gt):
        if g is None or not ctx.bcast:
            return g, None, None, None, None
        g = f32(g)
        B, N, H = g.shape
        out = torch.empty((B, H), device=g.device, dtype=torch.float32)
        call("xggm_sum_nodes", ptr(g), ptr(out), B, N, H)
        return out, None, None, None, None


NOISE_SITE_BASE = 1 << 40     # Philox subsequences of the in-kernel Gaussian draws (kept apart from the dropout sites)


def feat_noise_philox(feats, shape, sigma, planes):
    philox = (torch.initial_seed(), NOISE_SITE_BASE + _next_sites(1), _drop.epoch)
    return _FeatNoisePhilox.apply(feats, tuple(shape), sigma, planes, philox)


class _ScoreMse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, score, target, sigma):
        score, target = f32(score, "score"), f32(target, "target")
        loss = torch.empty(1, device=score.device, dtype=torch.float32)
        call("xggm_score_mse_fwd", ptr(score), ptr(target), float(sigma), ptr(loss), score.numel())
        ctx.save_for_backward(score, target)
        ctx.sigma = float(sigma)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        score, target = ctx.saved_tensors
        g = f32(g).reshape(1)
        gs = torch.empty_like(score)
        call("xggm_score_mse_bwd", ptr(score), ptr(target), ptr(g), ctx.sigma, ptr(gs), score.numel())
        return gs, (-gs if ctx.needs_input_grad[1] else None), None


class _SymKl(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        x, y = f32(x, "x"), f32(y, "y")
        if x.shape != y.shape:
            raise RuntimeError("xggm_b200.sym_kl: shape mismatch")
        C = x.shape[-1]
        R = x.numel() // C
        loss = torch.empty(1, device=x.device, dtype=torch.float32)
        call("xggm_sym_kl_fwd", ptr(x), ptr(y), ptr(loss), R, C)
        ctx.save_for_backward(x, y)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        C = x.shape[-1]
        R = x.numel() // C
        g = f32(g).reshape(1)
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        call("xggm_sym_kl_bwd", ptr(x), ptr(y), ptr(g), ptr(gx), ptr(gy), R, C)
        return gx, gy


class _FuseReadout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xp, nodes):
        xp, nodes = f32(xp, "x"), f32(nodes, "nodes")
        B, N, H = nodes.shape
        out = torch.empty((B, 2 * H), device=xp.device, dtype=torch.float32)
        call("xggm_fuse_readout_fwd", ptr(xp), ptr(nodes), ptr(out), B, N, H)
        ctx.save_for_backward(out)
        ctx.shape = (B, N, H)
        return out

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        B, N, H = ctx.shape
        g = f32(g)
        gxp = torch.empty((B, H), device=g.device, dtype=torch.float32)
        gn = torch.empty((B, N, H), device=g.device, dtype=torch.float32)
        call("xggm_fuse_readout_bwd", ptr(g), ptr(out), ptr(gxp), ptr(gn), B, N, H, 0)
        return gxp, gn


class _NodeTail(torch.autograd.Function):
    """Node-branch tail (vqacpv2.py:236-246): weighted KL + score-matching loss and the fusion_fc input."""

    @staticmethod
    def forward(ctx, nodes, feat, target, xp, sigma, kl_w, sm_w):
        nodes, feat, target, xp = f32(nodes, "nodes"), f32(feat, "feat"), f32(target, "target"), f32(xp, "x")
        B, N, H = nodes.shape
        if feat.shape != nodes.shape or target.shape != nodes.shape or xp.shape != (B, H):
            raise RuntimeError("xggm_b200.node_tail: shape mismatch")
        loss = torch.empty(1, device=nodes.device, dtype=torch.float32)
        cat = torch.empty((B, 2 * H), device=nodes.device, dtype=torch.float32)
        call("xggm_node_tail_fwd", ptr(nodes), ptr(feat), ptr(target), ptr(xp), float(sigma), float(kl_w), float(sm_w),
             ptr(loss), ptr(cat), B, N, H)
        ctx.save_for_backward(nodes, feat, target, cat)
        ctx.cfg = (float(sigma), float(kl_w), float(sm_w))
        return loss.reshape(()), cat

    @staticmethod
    def backward(ctx, gloss, gcat):
        nodes, feat, target, cat = ctx.saved_tensors
        sigma, kl_w, sm_w = ctx.cfg
        B, N, H = nodes.shape
        gloss = (torch.zeros(1, device=nodes.device) if gloss is None else f32(gloss)).reshape(1)
        gcat = torch.zeros_like(cat) if gcat is None else f32(gcat)
        gn = torch.empty_like(nodes)
        gf = torch.empty_like(feat) if ctx.needs_input_grad[1] else None
        gxp = torch.empty((B, H), device=nodes.device, dtype=torch.float32)
        grow = torch.empty((B, H), device=nodes.device, dtype=torch.float32)
        call("xggm_node_tail_bwd", ptr(nodes), ptr(feat), ptr(target), ptr(cat), ptr(gloss), ptr(gcat), sigma, kl_w, sm_w,
             ptr(gn), ptr(gf), ptr(gxp), ptr(grow), B, N, H)
        # the score-matching target's own gradient (never needed by the trainers) is the negated SM part
        return gn, gf, None, gxp, None, None, None


def node_tail(nodes, feat, target, xp, sigma, kl_w, sm_w):
    """(kl_w * compute_kl_loss(nodes, feat) + sm_w * loss_func(nodes, target, sigma),
    cat[xp, tanh(mean_n nodes)]) in one pass over the node features per direction."""
    return _NodeTail.apply(nodes, feat, target, xp, sigma, kl_w, sm_w)


def fuse_readout(xp, nodes):
    """cat[x, tanh(mean_n nodes)] (src/vqa/vqacpv2.py:216-218)."""
    return _FuseReadout.apply(xp, nodes)


class _Sigmoid(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = f32(x, "x")
        y = torch.empty_like(x)
        call("xggm_sigmoid_fwd", ptr(x), ptr(y), x.numel())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        g = f32(g)
        gx = torch.empty_like(y)
        call("xggm_sigmoid_bwd", ptr(g), ptr(y), ptr(gx), y.numel())
        return gx


def sigmoid(x):
    return _Sigmoid.apply(x)


# ---------------------------------------------------------------------------
# dropout keep-masks (Philox, generated on the device by the library)
# ---------------------------------------------------------------------------
class _DropState:
    site = 0       # next Philox subsequence id (one per dropout site)
    epoch = None   # device int64 counter while a CUDA-graph-captured step is being built / replayed


_drop = _DropState()
_mask_feed = []


def _next_sites(n):
    s = _drop.site
    _drop.site += n
    return s


class dropout_epoch:
    """Context for CUDA-graph capture: dropout sites are numbered from ``base`` on every entry (so the
    captured kernels and their replays agree) and every site adds ``epoch`` -- a device int64 tensor
    the captured step increments -- to its Philox subsequence, so each replay draws new masks."""

    def __init__(self, epoch, base=0):
        self.epoch, self.base = epoch, base

    def __enter__(self):
        self.old = (_drop.site, _drop.epoch)
        _drop.site, _drop.epoch = self.base, self.epoch
        return self

    def __exit__(self, *exc):
        _drop.site, _drop.epoch = self.old
        return False


def keep_mask(shape, p, device):
    """uint8 keep-mask: injected (parity runs, see inject_keep_masks) or Philox-generated
    keyed by (torch.initial_seed(), call counter) -- deterministic under torch.manual_seed."""
    if _mask_feed:
        m = _mask_feed.pop(0)
        if tuple(m.shape) != tuple(shape):
            raise RuntimeError(f"xggm_b200: injected keep-mask has shape {tuple(m.shape)}, need {tuple(shape)}")
        return _u8(m.to(device))
    m = torch.empty(shape, device=device, dtype=torch.uint8)
    _lib.check_device(m)
    call("xggm_keep_mask", ptr(m), m.numel(), float(p), torch.initial_seed() & (2 ** 64 - 1), _next_sites(1),
         ptr(_drop.epoch))
    return m


class inject_keep_masks:
    """Context manager: the next dropout sites consume these masks in call order
    (what oracle/make_golden.py does to the reference's F.dropout)."""

    def __init__(self, masks):
        self.masks = list(masks)

    def __enter__(self):
        _mask_feed.extend(self.masks)
        return self

    def __exit__(self, *exc):
        left = len(_mask_feed)
        _mask_feed.clear()
        if exc[0] is None and left:
            raise RuntimeError(f"xggm_b200: {left} injected keep-masks were not consumed")
        return False
