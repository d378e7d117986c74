"""torch.autograd Functions over the update-path kernels (include/crowdnav_b200.h "PPO update path"):

  * `linear_tc(x, w, b, act)`  -- Y = act(X W^T + b) with forward, data gradient and weight gradient on the tcgen05
    3xFP16 GEMM (cn_update_linear_fwd / _bwd), fp32-equivalent accuracy (dynamic power-of-two operand scales);
  * `hh_attention_rows(qkv, row_start, row_env)` -- the nn.MultiheadAttention core over COMPACTED rows (only valid
    humans), forward with soft-max statistics and a two-pass backward (cn_update_attn_fwd / _bwd).

Both exist only on CUDA tensors (the CPU path of Policy.evaluate_actions keeps plain torch ops, which is also what
the parity tests compare against).  Everything is enqueued on the current stream; no host synchronisation."""
import ctypes as C

import torch

from . import _capi

_ws = {}          # device index -> growing uint8 workspace


def _workspace(dev, nbytes):
    w = _ws.get(dev.index)
    if w is None or w.numel() < nbytes:
        w = _ws[dev.index] = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=dev)
    return w


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class _LinearTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, act):
        lib = _capi.load_library()
        x = x.contiguous()
        w = w.contiguous()
        M, K = x.shape
        N = w.shape[0]
        dev = x.device
        y = torch.empty(M, N, device=dev, dtype=torch.float32)
        saved = torch.empty(lib.cn_update_linear_saved_bytes(M, K), dtype=torch.uint8, device=dev)
        nws = lib.cn_update_linear_ws_bytes(M, N, K)
        ws = _workspace(dev, nws)
        bb = b.contiguous() if b is not None else None
        with torch.cuda.device(dev):
            _capi.check(lib, lib.cn_update_linear_fwd(_p(x), _p(w), _p(bb), _p(y), _p(saved), _p(ws), ws.numel(), M, N, K,
                                                      int(act), dev.index, _stream(dev)), "cn_update_linear_fwd")
        ctx.act, ctx.dims, ctx.has_bias = int(act), (M, N, K), b is not None
        ctx.save_for_backward(saved, w, y if act else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _capi.load_library()
        saved, w, y = ctx.saved_tensors
        M, N, K = ctx.dims
        dev = dy.device
        dy = dy.contiguous()
        dx = torch.empty(M, K, device=dev, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        dw = torch.empty(N, K, device=dev, dtype=torch.float32)
        db = torch.empty(N, device=dev, dtype=torch.float32) if ctx.has_bias else None
        ws = _workspace(dev, lib.cn_update_linear_ws_bytes(M, N, K))
        with torch.cuda.device(dev):
            _capi.check(lib, lib.cn_update_linear_bwd(_p(dy), _p(y), _p(saved), _p(w), _p(dx), _p(dw), _p(db), _p(ws),
                                                      ws.numel(), M, N, K, ctx.act, dev.index, _stream(dev)),
                        "cn_update_linear_bwd")
        return dx, dw, db, None


def linear_tc(x, w, b=None, act=0):
    """act(x @ w.T + b) for x [M, K], w [N, K] (N, K multiples of 64), act 0 = none / 1 = ReLU."""
    return _LinearTC.apply(x, w, b, act)


def linear_tc_supported(x, w):
    return x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[0] > 0 and w.shape[0] % 64 == 0 \
        and w.shape[1] % 64 == 0


class _HHAttentionRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, row_start, row_env):
        lib = _capi.load_library()
        qkv = qkv.contiguous()
        Mc = qkv.shape[0]
        dev = qkv.device
        out = torch.empty(Mc, 512, device=dev, dtype=torch.float32)
        stats = torch.empty(Mc, 16, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _capi.check(lib, lib.cn_update_attn_fwd(_p(qkv), _p(row_start), _p(row_env), Mc, _p(out), _p(stats), dev.index,
                                                    _stream(dev)), "cn_update_attn_fwd")
        ctx.save_for_backward(qkv, out, stats, row_start, row_env)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _capi.load_library()
        qkv, out, stats, row_start, row_env = ctx.saved_tensors
        Mc = qkv.shape[0]
        dev = qkv.device
        dout = dout.contiguous()
        dqkv = torch.empty_like(qkv)
        delta = torch.empty(Mc, 8, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _capi.check(lib, lib.cn_update_attn_bwd(_p(qkv), _p(out), _p(dout), _p(stats), _p(row_start), _p(row_env), Mc,
                                                    _p(dqkv), _p(delta), dev.index, _stream(dev)), "cn_update_attn_bwd")
        return dqkv, None, None


def hh_attention_rows(qkv, row_start, row_env):
    """qkv [Mc, 1536] fp32 CUDA, row_start int32 [B + 1], row_env int32 [Mc] -> [Mc, 512]."""
    return _HHAttentionRows.apply(qkv, row_start, row_env)


class _GruSeq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gi, h0, masks, whh, bhh):
        lib = _capi.load_library()
        T, N = gi.shape[0], gi.shape[1]
        dev = gi.device
        gi, h0, masks, whh, bhh = [t.contiguous() for t in (gi, h0, masks, whh, bhh)]
        out = torch.empty(T, N, 128, device=dev, dtype=torch.float32)
        saved = torch.empty(T, N, 512, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _capi.check(lib, lib.cn_update_gru_fwd(_p(gi), _p(h0), _p(masks), _p(whh), _p(bhh), T, N, _p(out), _p(saved), dev.index,
                                                   _stream(dev)), "cn_update_gru_fwd")
        ctx.save_for_backward(out, h0, masks, saved, whh)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _capi.load_library()
        out, h0, masks, saved, whh = ctx.saved_tensors
        T, N = out.shape[0], out.shape[1]
        dev = out.device
        dout = dout.contiguous()
        dgi = torch.empty(T, N, 384, device=dev, dtype=torch.float32)
        dghn = torch.empty(T, N, 128, device=dev, dtype=torch.float32)
        dh0 = torch.empty(N, 128, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _capi.check(lib, lib.cn_update_gru_bwd(_p(dout), None, _p(out), _p(h0), _p(masks), _p(saved), _p(whh), T, N, _p(dgi),
                                                   _p(dghn), _p(dh0), dev.index, _stream(dev)), "cn_update_gru_bwd")
        # recurrent weight / bias gradients: one GEMM over all steps (gh = hm W_hh^T + b_hh, hm = masked previous state)
        dgh = torch.cat([dgi[..., :256], dghn], -1).reshape(T * N, 384)
        hm = (torch.cat([h0.unsqueeze(0), out[:-1]], 0) * masks.unsqueeze(-1)).reshape(T * N, 128)
        return dgi, dh0, None, dgh.t() @ hm, dgh.sum(0)


def gru_sequence(gi, h0, masks, whh, bhh):
    """GRU cell over T steps with done-mask resets: gi [T, N, 384] (= x W_ih^T + b_ih), h0 [N, 128], masks [T, N] ->
    hidden state after every step [T, N, 128]."""
    return _GruSeq.apply(gi, h0, masks, whh, bhh)
