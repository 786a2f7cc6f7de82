"""Tensor-level wrappers over the C ABI (include/vtc_b200.h).

PyTorch is plumbing here: it owns device memory and the stream.  Every function takes CUDA
tensors, hands raw pointers + the current stream to libvtc_b200.so and returns CUDA tensors;
nothing synchronises.  There is no CPU path: a CPU tensor or a missing library raises.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Iterable, Optional, Sequence, Tuple

import torch

from . import _ffi
from ._ffi import VtcError

_PREC = {"exact": _ffi.PREC_EXACT, "fp32": _ffi.PREC_EXACT, "bf16": _ffi.PREC_BF16,
         "brute": _ffi.PREC_BRUTE}
_METRIC = {"dot": _ffi.METRIC_DOT, "l2": _ffi.METRIC_L2}


def _prec(p) -> int:
    if isinstance(p, int):
        return p
    try:
        return _PREC[p]
    except KeyError:
        raise ValueError(f"precision must be one of {sorted(_PREC)}, got {p!r}")


def _metric(m) -> int:
    if isinstance(m, int):
        return m
    try:
        return _METRIC[m]
    except KeyError:
        raise ValueError(f"metric must be 'l2' or 'dot', got {m!r}")


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _ffi.F32
    if t.dtype == torch.bfloat16:
        return _ffi.BF16
    raise TypeError(f"embeddings must be float32 or bfloat16, got {t.dtype}")


def _req_cuda(*ts: torch.Tensor) -> torch.device:
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise VtcError("vtc_b200 ops need CUDA tensors (there is no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise VtcError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


def _mat(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dim() != 2:
        raise ValueError(f"{name} must be 2-D [rows, D], got shape {tuple(t.shape)}")
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(dev: torch.device):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


# workspace cache: one growing byte buffer per (device, stream); use is stream-ordered
_ws: Dict[Tuple[int, int], torch.Tensor] = {}


def _workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    key = (dev.index if dev.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(dev).cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=dev)
        _ws[key] = buf
    return buf


def release_workspaces() -> None:
    _ws.clear()


def _ws_bytes(op: int, N: int, M: int, D: int, prec: int) -> int:
    n = _ffi.load().vtc_workspace_bytes(op, N, M, D, prec)
    if n == 0:
        raise VtcError("vtc_workspace_bytes rejected the arguments")
    return int(n)


# ------------------------------------------------------------------------------------------ H1
def normalize(x: torch.Tensor) -> torch.Tensor:
    """x / x.norm(dim=-1, keepdim=True) (model/model.py:26-27), any leading shape."""
    dev = _req_cuda(x)
    shp = x.shape
    x2 = x.reshape(-1, shp[-1]).contiguous()
    y = torch.empty_like(x2)
    with torch.cuda.device(dev):
        _ffi.check(_ffi.load().vtc_normalize(_ptr(x2), x2.shape[0], x2.shape[1], x2.shape[1],
                                             _dtype_code(x2), _ptr(y), x2.shape[1], _stream(dev)),
                   "vtc_normalize")
    return y.reshape(shp)


def row_norms(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(1/||x_r||, ||x_r||^2) per row, fp32: the norms behind model/model.py:26-27 and the
    ||x||^2 term of faiss' L2 distance (model/metric.py:144-146)."""
    dev = _req_cuda(x)
    x = _mat(x, "x")
    inv = torch.empty(x.shape[0], dtype=torch.float32, device=dev)
    sq = torch.empty_like(inv)
    with torch.cuda.device(dev):
        _ffi.check(_ffi.load().vtc_row_norms(_ptr(x), x.shape[0], x.shape[1], x.shape[1],
                                             _dtype_code(x), _ptr(inv), _ptr(sq), _stream(dev)),
                   "vtc_row_norms")
    return inv, sq


# ------------------------------------------------------------------------------------------ H2
def _scale_tensor(scale, dev) -> torch.Tensor:
    if isinstance(scale, torch.Tensor):
        return scale.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
    return torch.full((1,), float(scale), dtype=torch.float32, device=dev)


def sim_matrix(a: torch.Tensor, b: torch.Tensor, scale=1.0, precision="exact") -> torch.Tensor:
    """(scale * a) @ b.t() materialised as fp32 [N, M] (model/model.py:369)."""
    dev = _req_cuda(a, b)
    a, b = _mat(a, "a"), _mat(b, "b")
    if a.shape[1] != b.shape[1] or a.dtype != b.dtype:
        raise ValueError("a and b must share D and dtype")
    N, D = a.shape
    M = b.shape[0]
    prec = _prec(precision)
    sc = _scale_tensor(scale, dev)
    out = torch.empty(N, M, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(dev, _ws_bytes(_ffi.OP_SIM_MATRIX, N, M, D, prec))
        _ffi.check(_ffi.load().vtc_sim_matrix(_ptr(a), _ptr(b), N, M, D, _dtype_code(a), prec,
                                              _ptr(sc), _ptr(out), M, _ptr(ws), ws.numel(),
                                              _stream(dev)), "vtc_sim_matrix")
    return out


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
           residual: Optional[torch.Tensor] = None, act: int = 0, precision="exact") -> torch.Tensor:
    """act(x @ w.t() + bias) + residual on the tensor cores; x [rows, in], w [out, in], fp32.
    The CAM transformer's linears (in_proj / out_proj / c_fc / c_proj of clip.model.Transformer,
    structure per model/timesformer_clip_alt.py:43-67,112-124) and final_linear (model/model.py:161)."""
    dev = _req_cuda(x, w, bias, residual)
    x, w = _mat(x.float(), "x"), _mat(w.float(), "w")
    rows, in_f = x.shape
    out_f = w.shape[0]
    if w.shape[1] != in_f:
        raise ValueError("x and w disagree on in_features")
    if bias is not None:
        bias = bias.float().contiguous()
    if residual is not None:
        residual = residual.float().contiguous()
        if residual.shape != (rows, out_f):
            raise ValueError("residual must be [rows, out_features]")
    prec = _prec(precision)
    y = torch.empty(rows, out_f, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(dev, _ws_bytes(_ffi.OP_LINEAR, rows, out_f, in_f, prec))
        _ffi.check(_ffi.load().vtc_linear(_ptr(x), _ptr(w), _ptr(bias), _ptr(residual), rows, in_f,
                                          out_f, act, prec, _ptr(y), _ptr(ws), ws.numel(),
                                          _stream(dev)), "vtc_linear")
    return y


# ------------------------------------------------------------------------------------ R1 / R3
def rank_prepare(x: torch.Tensor, precision="exact", want_sq64: bool = True, want_qq: bool = True,
                 sq64_out: Optional[torch.Tensor] = None, qq_out: Optional[torch.Tensor] = None
                 ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Per-row quantities of a block of gallery / query rows, computed once and handed to every
    :func:`sim_rank` call that touches the block (`sq64=` for its gallery rows, `qq=` for its query
    rows): canonical fp64 ||x||^2 (what faiss' L2 distance adds per gallery row,
    model/metric.py:140-146) and an fp32 upper bound of ||x||^2 for the guard band.  The rows must be
    the canonical values: bf16 rows in the bf16 mode, fp32 rows in the exact mode."""
    dev = _req_cuda(x)
    x = _mat(x, "x")
    rows, D = x.shape
    sq64, qq = sq64_out, qq_out  # optional caller-owned outputs (e.g. slices of per-gallery arrays)
    for t, dt in ((sq64, torch.float64), (qq, torch.float32)):
        if t is not None and (t.dtype != dt or t.shape != (rows,) or not t.is_contiguous()
                              or t.device != dev):
            raise ValueError("rank_prepare outputs must be contiguous [rows] tensors on x's device")
    if sq64 is None and want_sq64:
        sq64 = torch.empty(rows, dtype=torch.float64, device=dev)
    if qq is None and want_qq:
        qq = torch.empty(rows, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _ffi.check(_ffi.load().vtc_rank_prepare(_ptr(x), rows, D, _dtype_code(x), _prec(precision),
                                                _ptr(sq64), _ptr(qq), _stream(dev)),
                   "vtc_rank_prepare")
    return sq64, qq


def sim_rank(q: torch.Tensor, g: torch.Tensor, gt: Optional[torch.Tensor] = None,
             row_offset: int = 0, col_offset: int = 0, metric="l2", precision="exact",
             gt_score: Optional[torch.Tensor] = None, rank0: Optional[torch.Tensor] = None,
             accumulate: bool = False, sq64: Optional[torch.Tensor] = None,
             qq: Optional[torch.Tensor] = None, sq64_out: Optional[torch.Tensor] = None,
             qq_out: Optional[torch.Tensor] = None, gt_score_out: Optional[torch.Tensor] = None
             ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fused similarity + rank of ground truth: replaces faiss GpuIndexFlatL2.add/search + the Python
    hit loop of RecallAtK.compute (model/metric.py:140-160).  Returns (rank0 int32 [N], gt_score
    fp64 [N]).

    rank0 is NOT finalised (NaN ground truths still hold their partial count): call
    :func:`rank_finalize` once every gallery chunk has been accumulated.

    Chunked evaluations cache three per-row quantities across calls (vtc_sim_rank_prepared): the
    canonical ||x||^2 of these gallery rows (`sq64` in, or `sq64_out` to have this call compute and
    store them), an upper bound of ||q||^2 of these query rows (`qq` / `qq_out`) and d(t,gt)
    (`gt_score` / `gt_score_out`).  Same result as the plain call; rows that are handed in with all
    three cached are not walked again."""
    dev = _req_cuda(q, g, gt, gt_score, rank0, sq64, qq, sq64_out, qq_out, gt_score_out)
    q, g = _mat(q, "q"), _mat(g, "g")
    if q.shape[1] != g.shape[1] or q.dtype != g.dtype:
        raise ValueError("queries and gallery must share D and dtype")
    N, D = q.shape
    M = g.shape[0]
    prec, met = _prec(precision), _metric(metric)
    if gt is not None:
        gt = gt.to(torch.int64).contiguous()
        if gt.shape != (N,):
            raise ValueError("gt must be [N]")
    if rank0 is None:
        rank0 = torch.empty(N, dtype=torch.int32, device=dev)  # every entry is written
        accumulate = False
    elif rank0.dtype != torch.int32 or rank0.shape != (N,) or not rank0.is_contiguous():
        raise ValueError("rank0 must be a contiguous int32 [N] tensor")

    def vec(t, dt, n, name):
        if t is not None and (t.dtype != dt or t.shape != (n,) or not t.is_contiguous()):
            raise ValueError(f"{name} must be a contiguous {dt} [{n}] tensor")
        return t

    gt_score = vec(gt_score, torch.float64, N, "gt_score")
    gs_out = vec(gt_score_out, torch.float64, N, "gt_score_out")
    if gt_score is None and gs_out is None:
        gs_out = torch.empty(N, dtype=torch.float64, device=dev)
    cached = any(t is not None for t in (sq64, qq, sq64_out, qq_out))
    if cached and prec != _ffi.PREC_BRUTE:
        sq64 = vec(sq64, torch.float64, M, "sq64")
        qq = vec(qq, torch.float32, N, "qq")
        sq64_out = vec(sq64_out, torch.float64, M, "sq64_out")
        qq_out = vec(qq_out, torch.float32, N, "qq_out")
        if sq64 is None and sq64_out is None:
            sq64_out = torch.empty(M, dtype=torch.float64, device=dev)
        if qq is None and qq_out is None:
            qq_out = torch.empty(N, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _workspace(dev, _ws_bytes(_ffi.OP_SIM_RANK, N, M, D, prec))
            _ffi.check(_ffi.load().vtc_sim_rank_prepared(
                _ptr(q), _ptr(g), N, M, D, _dtype_code(q), _ptr(gt), row_offset, col_offset, met,
                prec, _ptr(gt_score), _ptr(None if gt_score is not None else gs_out),
                _ptr(sq64), _ptr(None if sq64 is not None else sq64_out),
                _ptr(qq), _ptr(None if qq is not None else qq_out),
                1 if accumulate else 0, _ptr(rank0), _ptr(ws), ws.numel(), _stream(dev)),
                "vtc_sim_rank_prepared")
        return rank0, (gt_score if gt_score is not None else gs_out)
    with torch.cuda.device(dev):
        ws = _workspace(dev, _ws_bytes(_ffi.OP_SIM_RANK, N, M, D, prec))
        _ffi.check(_ffi.load().vtc_sim_rank(
            _ptr(q), _ptr(g), N, M, D, _dtype_code(q), _ptr(gt), row_offset, col_offset, met, prec,
            _ptr(gt_score), _ptr(gs_out), 1 if accumulate else 0, _ptr(rank0), _ptr(ws), ws.numel(),
            _stream(dev)), "vtc_sim_rank")
    return rank0, (gt_score if gt_score is not None else gs_out)


def rank_eval(q: torch.Tensor, g: torch.Tensor, k_vals: Sequence[int],
              gt: Optional[torch.Tensor] = None, metric="l2", precision="exact",
              want_medr: bool = True) -> Dict[str, Optional[torch.Tensor]]:
    """The whole of RecallAtK.compute (model/metric.py:137-161) for device-resident embeddings in
    one library call -- memset + 5 launches (prologue, tensor-core pass, re-check, commit, median
    select): similarity + rank of ground truth over the full gallery,
    NaN ground truths -> rank M, hit counts for every k, median rank.  Returns
    {"rank0": int32 [N], "hits": int64 [nk], "medr": fp64 [1] | None, "gt_score": fp64 [N]}."""
    k_vals = [int(k) for k in k_vals]
    if len(k_vals) > 8:  # the fused finalisation counts up to 8 thresholds (the reference uses 2-3)
        rank0, gts = sim_rank(q, g, gt=gt, metric=metric, precision=precision)
        hits, medr = rank_finalize(rank0, gts, g.shape[0], k_vals, want_medr)
        return {"rank0": rank0, "hits": hits, "medr": medr, "gt_score": gts}
    dev = _req_cuda(q, g, gt)
    q, g = _mat(q, "q"), _mat(g, "g")
    if q.shape[1] != g.shape[1] or q.dtype != g.dtype:
        raise ValueError("queries and gallery must share D and dtype")
    N, D = q.shape
    M = g.shape[0]
    prec, met = _prec(precision), _metric(metric)
    if gt is not None:
        gt = gt.to(torch.int64).contiguous()
        if gt.shape != (N,):
            raise ValueError("gt must be [N]")
    rank0 = torch.empty(N, dtype=torch.int32, device=dev)
    gts = torch.empty(N, dtype=torch.float64, device=dev)
    hits = torch.empty(max(1, len(k_vals)), dtype=torch.int64, device=dev)
    medr = torch.empty(1, dtype=torch.float64, device=dev) if want_medr else None
    karr = (ctypes.c_int * max(1, len(k_vals)))(*k_vals)
    with torch.cuda.device(dev):
        ws = _workspace(dev, _ws_bytes(_ffi.OP_SIM_RANK, N, M, D, prec))
        _ffi.check(_ffi.load().vtc_rank_eval(
            _ptr(q), _ptr(g), N, M, D, _dtype_code(q), _ptr(gt), met, prec, karr, len(k_vals),
            _ptr(rank0), _ptr(hits), _ptr(medr), _ptr(gts), _ptr(ws), ws.numel(), _stream(dev)),
            "vtc_rank_eval")
    return {"rank0": rank0, "hits": hits[:len(k_vals)], "medr": medr, "gt_score": gts}


def gt_scores(q: torch.Tensor, g: torch.Tensor, gt: Optional[torch.Tensor] = None,
              row_offset: int = 0, col_offset: int = 0, metric="l2", precision="exact"
              ) -> torch.Tensor:
    """fp64-sequential d(t, gt(t)); NaN where gt lies outside this gallery chunk."""
    dev = _req_cuda(q, g, gt)
    q, g = _mat(q, "q"), _mat(g, "g")
    N, D = q.shape
    M = g.shape[0]
    prec, met = _prec(precision), _metric(metric)
    if gt is not None:
        gt = gt.to(torch.int64).contiguous()
    out = torch.empty(N, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(dev, _ws_bytes(_ffi.OP_GT_SCORES, N, M, D, prec))
        _ffi.check(_ffi.load().vtc_gt_scores(_ptr(q), _ptr(g), N, M, D, _dtype_code(q), _ptr(gt),
                                             row_offset, col_offset, met, prec, _ptr(out), _ptr(ws),
                                             ws.numel(), _stream(dev)), "vtc_gt_scores")
    return out


def rank_finalize(rank0: torch.Tensor, gt_score: Optional[torch.Tensor], M_total: int,
                  k_vals: Sequence[int], want_medr: bool = True
                  ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """In place: rank0 = M_total where gt_score is NaN.  Returns (hits int64 [nk], medr fp64 [1]):
    the hit counts behind recall@k (model/metric.py:148-160) and MedR."""
    dev = _req_cuda(rank0, gt_score)
    k_vals = [int(k) for k in k_vals]
    if len(k_vals) > 8:
        # the entry point takes up to 8 thresholds per call (the reference's configs use 2-3,
        # model/metric.py:104); more are counted 8 at a time -- the NaN -> M_total pass is idempotent
        parts = [rank_finalize(rank0, gt_score, M_total, k_vals[i:i + 8], want_medr and i == 0)
                 for i in range(0, len(k_vals), 8)]
        return torch.cat([h for h, _ in parts]), parts[0][1]
    hits = torch.zeros(max(1, len(k_vals)), dtype=torch.int64, device=dev)
    medr = torch.empty(1, dtype=torch.float64, device=dev) if want_medr else None
    karr = (ctypes.c_int * max(1, len(k_vals)))(*k_vals)
    with torch.cuda.device(dev):
        hist = _workspace(dev, (3 * 65536 + 8) * 4)
        _ffi.check(_ffi.load().vtc_rank_finalize(_ptr(rank0), _ptr(gt_score), rank0.shape[0],
                                                 int(M_total), karr, len(k_vals), _ptr(hits),
                                                 _ptr(medr), _ptr(hist), hist.numel(),
                                                 _stream(dev)), "vtc_rank_finalize")
    return hits[:len(k_vals)], medr


# ------------------------------------------------------------------------------------------ K7
def sim_topk(q: torch.Tensor, g: torch.Tensor, k: int, metric="l2", precision="exact",
             col_offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fused similarity + streaming top-k: (vals fp32 [N,k] ascending, idx int64 [N,k]) -- what
    faiss_search_index.search(features_b, max(k_vals) + 1) returns (model/metric.py:144-146)."""
    dev = _req_cuda(q, g)
    q, g = _mat(q, "q"), _mat(g, "g")
    if q.shape[1] != g.shape[1] or q.dtype != g.dtype:
        raise ValueError("queries and gallery must share D and dtype")
    N, D = q.shape
    M = g.shape[0]
    prec, met = _prec(precision), _metric(metric)
    vals = torch.empty(N, k, dtype=torch.float32, device=dev)
    idx = torch.empty(N, k, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        ws = _workspace(dev, _ws_bytes(_ffi.OP_SIM_TOPK, N, M, D, prec))
        _ffi.check(_ffi.load().vtc_sim_topk(_ptr(q), _ptr(g), N, M, D, _dtype_code(q), met, prec, k,
                                            col_offset, _ptr(vals), _ptr(idx), _ptr(ws), ws.numel(),
                                            _stream(dev)), "vtc_sim_topk")
    return vals, idx


def topk_merge(vals: torch.Tensor, idx: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge [parts, N, k] sorted candidate lists into the k best per row."""
    dev = _req_cuda(vals, idx)
    vals = vals.float().contiguous()
    idx = idx.to(torch.int64).contiguous()
    parts, N, k = vals.shape
    ov = torch.empty(N, k, dtype=torch.float32, device=dev)
    oi = torch.empty(N, k, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _ffi.check(_ffi.load().vtc_topk_merge(_ptr(vals), _ptr(idx), parts, N, k, _ptr(ov), _ptr(oi),
                                              _stream(dev)), "vtc_topk_merge")
    return ov, oi


# ------------------------------------------------------------------------------------- H2 + H3
def infonce_fwd(a: torch.Tensor, b: torch.Tensor, scale, precision="exact"
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Fused symmetric InfoNCE forward, 0.5 * (CE(sim, arange) + CE(sim.t(), arange)) with
    sim = scale * a @ b.t() never materialised (model/loss.py:18-22, model/model.py:369):
    (loss [1], row_lse [n], col_lse [n], diag [n])."""
    dev = _req_cuda(a, b)
    a, b = _mat(a, "a"), _mat(b, "b")
    if a.shape != b.shape or a.dtype != b.dtype:
        raise ValueError("clip_loss needs a square similarity: a and b must have the same shape")
    n, D = a.shape
    prec = _prec(precision)
    sc = _scale_tensor(scale, dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    row = torch.empty(n, dtype=torch.float32, device=dev)
    col = torch.empty_like(row)
    diag = torch.empty_like(row)
    with torch.cuda.device(dev):
        ws = _workspace(dev, _ws_bytes(_ffi.OP_INFONCE_FWD, n, n, D, prec))
        _ffi.check(_ffi.load().vtc_infonce_fwd(_ptr(a), _ptr(b), n, D, _dtype_code(a), prec,
                                               _ptr(sc), _ptr(loss), _ptr(row), _ptr(col),
                                               _ptr(diag), _ptr(ws), ws.numel(), _stream(dev)),
                   "vtc_infonce_fwd")
    return loss, row, col, diag


def infonce_bwd(a: torch.Tensor, b: torch.Tensor, scale, row_lse: torch.Tensor,
                col_lse: torch.Tensor, grad_loss: torch.Tensor, precision="exact"
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Backward of infonce_fwd w.r.t. both feature matrices and the logit scale (the autograd of
    model/loss.py:18-22 through model/model.py:369): (dA [n,D], dB [n,D], dscale [1]) fp32.
    `precision` must be the forward's: the logits are recomputed from the same operands."""
    dev = _req_cuda(a, b, row_lse, col_lse, grad_loss)
    a, b = _mat(a, "a"), _mat(b, "b")
    n, D = a.shape
    sc = _scale_tensor(scale, dev)
    g = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
    dA = torch.empty(n, D, dtype=torch.float32, device=dev)
    dB = torch.empty_like(dA)
    ds = torch.empty(1, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        prec = _prec(precision)
        ws = _workspace(dev, _ws_bytes(_ffi.OP_INFONCE_BWD, n, n, D, prec))
        _ffi.check(_ffi.load().vtc_infonce_bwd(_ptr(a), _ptr(b), n, D, _dtype_code(a), prec, _ptr(sc),
                                               _ptr(row_lse), _ptr(col_lse), _ptr(g), _ptr(dA),
                                               _ptr(dB), _ptr(ds), _ptr(ws), ws.numel(),
                                               _stream(dev)), "vtc_infonce_bwd")
    return dA, dB, ds


def infonce_dense_fwd(sim: torch.Tensor):
    """clip_loss of a materialised fp32 [n, n] sim: (loss [1], row_lse, col_lse, diag)."""
    dev = _req_cuda(sim)
    sim = _mat(sim, "sim")
    n = sim.shape[0]
    if sim.shape[1] != n or sim.dtype != torch.float32:
        raise ValueError("sim must be a square fp32 matrix")
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    row, col, diag = (torch.empty(n, dtype=torch.float32, device=dev) for _ in range(3))
    with torch.cuda.device(dev):
        _ffi.check(_ffi.load().vtc_infonce_dense_fwd(_ptr(sim), n, n, _ptr(loss), _ptr(row), _ptr(col),
                                                     _ptr(diag), _stream(dev)), "vtc_infonce_dense_fwd")
    return loss, row, col, diag


def infonce_dense_bwd(sim: torch.Tensor, row_lse: torch.Tensor, col_lse: torch.Tensor,
                      grad_loss: torch.Tensor) -> torch.Tensor:
    dev = _req_cuda(sim, row_lse, col_lse, grad_loss)
    sim = _mat(sim, "sim")
    n = sim.shape[0]
    g = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
    dsim = torch.empty_like(sim)
    with torch.cuda.device(dev):
        _ffi.check(_ffi.load().vtc_infonce_dense_bwd(_ptr(sim), n, n, _ptr(row_lse), _ptr(col_lse),
                                                     _ptr(g), _ptr(dsim), n, _stream(dev)),
                   "vtc_infonce_dense_bwd")
    return dsim


# ------------------------------------------------------------------------------------------ H4
def cam_stack_normalize(main: torch.Tensor, aux: torch.Tensor) -> torch.Tensor:
    """normalize(stack([main, *aux])) -> [1+nc, b, D] (model/model.py:150-151)."""
    dev = _req_cuda(main, aux)
    main = main.float().contiguous()
    aux = aux.float().contiguous()
    b, D = main.shape
    nc = aux.shape[0]
    if aux.shape[1:] != (b, D):
        raise ValueError("aux must be [nc, b, D]")
    X = torch.empty(nc + 1, b, D, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _ffi.check(_ffi.load().vtc_cam_stack_normalize(_ptr(main), _ptr(aux), nc + 1, b, D, _ptr(X),
                                                       _stream(dev)), "vtc_cam_stack_normalize")
    return X


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5
              ) -> torch.Tensor:
    """LayerNorm over the last dim in fp32 (ln_1 / ln_2 of clip.model.Transformer's blocks, the
    fp32-upcast LayerNorm of model/timesformer_clip_alt.py:22-33)."""
    dev = _req_cuda(x, gamma, beta)
    shp = x.shape
    x2 = x.float().reshape(-1, shp[-1]).contiguous()
    y = torch.empty_like(x2)
    with torch.cuda.device(dev):
        _ffi.check(_ffi.load().vtc_layernorm(_ptr(x2), _ptr(gamma.float().contiguous()),
                                             _ptr(beta.float().contiguous()), x2.shape[0],
                                             x2.shape[1], eps, _ptr(y), _stream(dev)),
                   "vtc_layernorm")
    return y.reshape(shp)


def cam_attn_core(qkv: torch.Tensor, heads: int) -> torch.Tensor:
    """softmax(q k^T / sqrt(hd)) v per (sample, head); qkv [L, b, 3D] -> [L, b, D]: the attention
    core of nn.MultiheadAttention inside clip.model.Transformer (model/model.py:155; structure per
    model/timesformer_clip_alt.py:43-67), seq-first, no mask."""
    dev = _req_cuda(qkv)
    qkv = qkv.float().contiguous()
    L, b, D3 = qkv.shape
    D = D3 // 3
    out = torch.empty(L, b, D, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _ffi.check(_ffi.load().vtc_cam_attn_core(_ptr(qkv), L, b, D, heads, _ptr(out),
                                                 _stream(dev)), "vtc_cam_attn_core")
    return out


def _res_act_args(res_act, dev):
    """(act, scale, shift, mul) -> ctypes-ready (int, float, tensor|None, tensor|None)."""
    if res_act is None:
        return _ffi.RESACT_NONE, 1.0, None, None
    act, scale, shift, mul = res_act
    f = lambda t: None if t is None else t.detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
    return int(act), float(scale), f(shift), f(mul)


def cam_readout(T: Optional[torch.Tensor], main: Optional[torch.Tensor], mode: int,
                res_in: Optional[torch.Tensor] = None, skip_mask: Optional[torch.Tensor] = None,
                res_act=None) -> torch.Tensor:
    """Read-out of the CAM in one kernel (model/model.py:156-161, 168-171, 199-203): AVG =
    normalize(mean_l normalize(T_l)), RESIDUAL_ONLY = a caller-computed residual (final_linear of
    token 0), then the residual activation of model/model.py:30-77, the random adapter skip and
    normalize(normalize(main) + res); UNIFORM = the averaging fusion of model/model.py:356-366."""
    dev = _req_cuda(T, main, res_in, skip_mask)
    if T is not None:
        T = T.float().contiguous()
        L, b, D = T.shape
    else:
        L = 1
        b, D = res_in.shape
        res_in = res_in.float().contiguous()
    if main is not None:
        main = main.float().contiguous()
    if skip_mask is not None:
        skip_mask = skip_mask.to(device=dev, dtype=torch.uint8).contiguous()
    out = torch.empty(b, D, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        act, scale, shift, mul = _res_act_args(res_act, dev)
        _ffi.check(_ffi.load().vtc_cam_readout(_ptr(T), _ptr(main), _ptr(res_in), _ptr(skip_mask),
                                               L, b, D, mode, act, scale, _ptr(shift), _ptr(mul),
                                               _ptr(out), _stream(dev)),
                   "vtc_cam_readout")
    return out


# --------------------------------------------------------------------- H4 in one call (prepared)
def linear_prepare(w: torch.Tensor, bias: Optional[torch.Tensor], precision="exact") -> torch.Tensor:
    """Weight [out, in] (+bias) -> opaque prepared buffer (bf16 gallery-side operand + padded bias)."""
    dev = _req_cuda(w, bias)
    w = w.detach().float().contiguous()
    out_f, in_f = w.shape
    prec = _prec(precision)
    lib = _ffi.load()
    nbytes = lib.vtc_linear_prepared_bytes(in_f, out_f, prec)
    if nbytes == 0:
        raise VtcError("vtc_linear_prepared_bytes rejected the arguments")
    buf = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
    b = None if bias is None else bias.detach().float().contiguous()
    with torch.cuda.device(dev):
        _ffi.check(lib.vtc_linear_prepare(_ptr(w), _ptr(b), in_f, out_f, prec, _ptr(buf), _stream(dev)),
                   "vtc_linear_prepare")
    return buf


def cam_forward(main: torch.Tensor, aux: torch.Tensor, layers, heads: int, readout_mode: int,
                final_linear: Optional[torch.Tensor] = None, skip_mask: Optional[torch.Tensor] = None,
                precision="exact", res_act=None) -> torch.Tensor:
    """PretrainedCLIPBase._adapt_feature (model/model.py:141-205) in one C call.

    `layers` is a ctypes array of _ffi.CamLayer built from prepared linears (see
    vtc_b200.model.model.CAMTransformer.prepared)."""
    dev = _req_cuda(main, aux, final_linear, skip_mask)
    main = main.float().contiguous()
    aux = aux.float().contiguous()
    b, D = main.shape
    nc = aux.shape[0]
    if aux.shape[1:] != (b, D):
        raise ValueError("aux must be [nc, b, D]")
    prec = _prec(precision)
    if skip_mask is not None:
        skip_mask = skip_mask.to(device=dev, dtype=torch.uint8).contiguous()
    out = torch.empty(b, D, dtype=torch.float32, device=dev)
    lib = _ffi.load()
    with torch.cuda.device(dev):
        ws = _workspace(dev, int(lib.vtc_cam_workspace_bytes(nc + 1, b, D, prec)))
        act, scale, shift, mul = _res_act_args(res_act, dev)
        _ffi.check(lib.vtc_cam_forward(_ptr(main), _ptr(aux), nc + 1, b, D, heads, len(layers), layers,
                                       readout_mode, _ptr(final_linear), _ptr(skip_mask), act, scale,
                                       _ptr(shift), _ptr(mul), prec, _ptr(out), _ptr(ws), ws.numel(),
                                       _stream(dev)),
                   "vtc_cam_forward")
    return out


# ------------------------------------------------------------------------------ CAM backward ops
def _call(name, *args):
    _ffi.check(getattr(_ffi.load(), name)(*args), name)


def normalize_bwd(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    """dX of y = x / |x| over the last dim."""
    dev = _req_cuda(x, dy)
    shp = x.shape
    x2 = x.float().reshape(-1, shp[-1]).contiguous()
    d2 = dy.float().reshape(-1, shp[-1]).contiguous()
    out = torch.empty_like(x2)
    with torch.cuda.device(dev):
        _call("vtc_normalize_bwd", _ptr(x2), _ptr(d2), x2.shape[0], x2.shape[1], _ptr(out), _stream(dev))
    return out.reshape(shp)


def transpose(x: torch.Tensor) -> torch.Tensor:
    """[R, C] fp32 -> [C, R] (materialised; operands of the backward GEMMs)."""
    dev = _req_cuda(x)
    x = _mat(x.float(), "x")
    R, C = x.shape
    out = torch.empty(C, R, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _call("vtc_transpose", _ptr(x), R, C, _ptr(out), _stream(dev))
    return out


def gelu_bwd(dF: torch.Tensor, U: torch.Tensor) -> torch.Tensor:
    dev = _req_cuda(dF, U)
    dF, U = dF.float().contiguous(), U.float().contiguous()
    out = torch.empty_like(dF)
    with torch.cuda.device(dev):
        _call("vtc_gelu_bwd", _ptr(dF), _ptr(U), dF.numel(), _ptr(out), _stream(dev))
    return out


def colsum(x: torch.Tensor) -> torch.Tensor:
    dev = _req_cuda(x)
    x = _mat(x.float(), "x")
    out = torch.empty(x.shape[1], dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _call("vtc_colsum", _ptr(x), x.shape[0], x.shape[1], _ptr(out), _stream(dev))
    return out


def bias_act(x: torch.Tensor, bias: Optional[torch.Tensor] = None,
             residual: Optional[torch.Tensor] = None, act: int = 0) -> torch.Tensor:
    dev = _req_cuda(x, bias, residual)
    x = _mat(x.float(), "x")
    out = torch.empty_like(x)
    b = None if bias is None else bias.float().contiguous()
    r = None if residual is None else residual.float().contiguous()
    with torch.cuda.device(dev):
        _call("vtc_bias_act", _ptr(x), _ptr(b), _ptr(r), x.shape[0], x.shape[1], act, _ptr(out),
              _stream(dev))
    return out


def layernorm_bwd(dY: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, eps: float = 1e-5,
                  dres: Optional[torch.Tensor] = None):
    """-> (dX [rows, D] (+ dres), dgamma [D], dbeta [D])."""
    dev = _req_cuda(dY, x, gamma, dres)
    dY, x = _mat(dY.float(), "dY"), _mat(x.float(), "x")
    rows, D = x.shape
    g = gamma.detach().float().contiguous()
    dres = None if dres is None else dres.float().contiguous()
    dX = torch.empty_like(x)
    dg = torch.empty(D, dtype=torch.float32, device=dev)
    db = torch.empty_like(dg)
    with torch.cuda.device(dev):
        _call("vtc_layernorm_bwd", _ptr(dY), _ptr(x), _ptr(g), rows, D, eps, _ptr(dres), _ptr(dX),
              _ptr(dg), _ptr(db), _stream(dev))
    return dX, dg, db


def cam_attn_core_bwd(qkv: torch.Tensor, dO: torch.Tensor, heads: int) -> torch.Tensor:
    dev = _req_cuda(qkv, dO)
    qkv, dO = qkv.float().contiguous(), dO.float().contiguous()
    L, b, D3 = qkv.shape
    out = torch.empty_like(qkv)
    with torch.cuda.device(dev):
        _call("vtc_cam_attn_core_bwd", _ptr(qkv), _ptr(dO), L, b, D3 // 3, heads, _ptr(out), _stream(dev))
    return out


def cam_stack_normalize_bwd(main: torch.Tensor, aux: torch.Tensor, dX: torch.Tensor):
    dev = _req_cuda(main, aux, dX)
    main, aux, dX = main.float().contiguous(), aux.float().contiguous(), dX.float().contiguous()
    b, D = main.shape
    L = aux.shape[0] + 1
    dmain = torch.empty_like(main)
    daux = torch.empty_like(aux)
    with torch.cuda.device(dev):
        _call("vtc_cam_stack_normalize_bwd", _ptr(main), _ptr(aux), _ptr(dX), L, b, D, _ptr(dmain),
              _ptr(daux), _stream(dev))
    return dmain, daux


def cam_readout_bwd(T: Optional[torch.Tensor], main: torch.Tensor, dout: torch.Tensor, mode: int,
                    res_in: Optional[torch.Tensor] = None, skip_mask: Optional[torch.Tensor] = None,
                    L: int = 1, res_act=None):
    """-> (dT [L,b,D] or None, dres [b,D] or None, dmain [b,D]).  `res_act` as in cam_readout."""
    dev = _req_cuda(T, main, dout, res_in, skip_mask)
    dout = dout.float().contiguous()
    if mode == _ffi.CAM_READOUT_UNIFORM:
        T = T.float().contiguous()
        L, b, D = T.shape
        dT = torch.empty_like(T)
        with torch.cuda.device(dev):
            _call("vtc_cam_readout_bwd", _ptr(T), None, None, None, _ptr(dout), L, b, D, mode,
                  _ffi.RESACT_NONE, 1.0, None, None, _ptr(dT), None, None, _stream(dev))
        return dT, None, None
    main = main.float().contiguous()
    b, D = main.shape
    dT = dres = None
    if mode == _ffi.CAM_READOUT_AVG:
        T = T.float().contiguous()
        L = T.shape[0]
        dT = torch.empty_like(T)
    else:
        res_in = res_in.float().contiguous()
        dres = torch.empty_like(res_in)
    if skip_mask is not None:
        skip_mask = skip_mask.to(device=dev, dtype=torch.uint8).contiguous()
    dmain = torch.empty_like(main)
    with torch.cuda.device(dev):
        act, scale, shift, mul = _res_act_args(res_act, dev)
        _call("vtc_cam_readout_bwd", _ptr(T), _ptr(main), _ptr(res_in), _ptr(skip_mask), _ptr(dout), L, b,
              D, mode, act, scale, _ptr(shift), _ptr(mul), _ptr(dT), _ptr(dres), _ptr(dmain),
              _stream(dev))
    return dT, dres, dmain


# ----------------------------------------------------------------------------------- tracing
def cam_backward(dout: torch.Tensor, main: torch.Tensor, aux: torch.Tensor, T: torch.Tensor,
                 res_in: Optional[torch.Tensor], layers, heads: int, readout_mode: int,
                 final_linear_t: Optional[torch.Tensor] = None,
                 skip_mask: Optional[torch.Tensor] = None, precision="exact", res_act=None,
                 want_dflw: bool = False):
    """The whole backward of `_adapt_feature` in one C call (vtc_cam_backward).  `layers` is a ctypes
    array of _ffi.CamLayerBwd whose gradient buffers the call fills.  Returns (dmain, daux, dflw)."""
    dev = _req_cuda(dout, main, aux, T)
    dout = dout.float().contiguous()
    b, D = main.shape
    L = aux.shape[0] + 1
    prec = _prec(precision)
    dmain = torch.empty(b, D, dtype=torch.float32, device=dev)
    daux = torch.empty_like(aux, dtype=torch.float32)
    dflw = torch.empty(D, D, dtype=torch.float32, device=dev) if want_dflw else None
    if skip_mask is not None:
        skip_mask = skip_mask.to(device=dev, dtype=torch.uint8).contiguous()
    lib = _ffi.load()
    with torch.cuda.device(dev):
        ws = _workspace(dev, int(lib.vtc_cam_backward_workspace_bytes(L, b, D, prec)))
        act, scale, shift, mul = _res_act_args(res_act, dev)
        _ffi.check(lib.vtc_cam_backward(_ptr(dout), _ptr(main), _ptr(aux), _ptr(T), _ptr(res_in),
                                        _ptr(skip_mask), L, b, D, heads, len(layers), layers,
                                        readout_mode, _ptr(final_linear_t), act, scale, _ptr(shift),
                                        _ptr(mul), prec, _ptr(dmain), _ptr(daux), _ptr(dflw), _ptr(ws),
                                        ws.numel(), _stream(dev)), "vtc_cam_backward")
    return dmain, daux, dflw


def _install_nvtx_ranges() -> None:
    """VTC_NVTX=1: wrap every public op in an NVTX range `vtc.<name>` so that ncu / nsys can filter
    and group by library call (`ncu --nvtx --nvtx-include "vtc.sim_rank/"`).  The reference's only
    timing around this path is a wall-clock print in RecallAtK.result (model/metric.py:167-185)."""
    import functools
    import types

    def wrap(fn):
        @functools.wraps(fn)
        def ranged(*args, **kwargs):
            torch.cuda.nvtx.range_push("vtc." + fn.__name__)
            try:
                return fn(*args, **kwargs)
            finally:
                torch.cuda.nvtx.range_pop()
        return ranged

    g = globals()
    for name, fn in list(g.items()):
        if (isinstance(fn, types.FunctionType) and not name.startswith("_")
                and fn.__module__ == __name__ and name != "release_workspaces"):
            g[name] = wrap(fn)


import os as _os  # noqa: E402

if _os.environ.get("VTC_NVTX"):
    _install_nvtx_ranges()
