"""TEST INFRASTRUCTURE ONLY - Python face of the CPU oracle (see oracle/lutgemv_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import
this module.  The product path (squeezellm_b200/) must never import it.

Contents
  * ctypes bindings to oracle/liblutgemv_oracle.so (C restatement, OpenMP);
  * `unpack_np` / `pack_np`: an independent vectorised numpy restatement of the packed format
    (squeezellm/quant.py:171-208, squeezellm/quant_cuda_kernel.cu:776-825,863-877) used to
    cross-check the C code;
  * `make_layer`: synthetic QuantLinearLUT buffers in the reference's exact on-disk format
    (squeezellm/quant.py:48-95), as SURVEY.md section 8(d) specifies them;
  * `cpu_dequant_matmul`: the "dequant-to-fp16 + torch.matmul" CPU baseline named by
    BASELINE.json (a restatement - the reference has no CPU compute path).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile the C oracle in-tree (idempotent)."""
    so = os.path.join(HERE, "liblutgemv_oracle.so")
    src = os.path.join(HERE, "lutgemv_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.sq_oracle_threads.restype = ctypes.c_int
    return _LIB


def _p(a, ct=ctypes.c_void_p):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ct)


def threads():
    return int(lib().sq_oracle_threads())


# ------------------------------------------------------------------------------------------------
# C-backed entry points
# ------------------------------------------------------------------------------------------------
def unpack(qweight, bits):
    """qweight int32 [K/32*bits, N] -> idx uint8 [K, N].  (C restatement.)"""
    qweight = np.ascontiguousarray(qweight, dtype=np.int32)
    rows, N = qweight.shape
    K = rows * 32 // bits
    idx = np.empty((K, N), dtype=np.uint8)
    lib().sq_unpack(ctypes.c_int(bits), _p(qweight), ctypes.c_int(K), ctypes.c_int(N), _p(idx))
    return idx


def pack(idx, bits):
    """idx uint8 [K, N] -> qweight int32 [K/32*bits, N].  (C restatement of pack2's loop.)"""
    idx = np.ascontiguousarray(idx, dtype=np.uint8)
    K, N = idx.shape
    assert K % 32 == 0
    q = np.empty((K // 32 * bits, N), dtype=np.int32)
    lib().sq_pack(ctypes.c_int(bits), _p(idx), ctypes.c_int(K), ctypes.c_int(N), _p(q))
    return q


def _prep(layer, vec, mul_init):
    bits, K, N = layer["bits"], layer["infeatures"], layer["outfeatures"]
    vec = np.ascontiguousarray(vec, dtype=np.float32).reshape(-1, K)
    batch = vec.shape[0]
    f32 = lambda k: None if layer.get(k) is None else np.ascontiguousarray(layer[k], dtype=np.float32)
    i32 = lambda k: None if layer.get(k) is None else np.ascontiguousarray(layer[k], dtype=np.int32)
    args = dict(qweight=i32("qweight"), lut=f32("lookup_table"), rows=i32("rows"), cols=i32("cols"),
                vals=f32("vals"), full_rows=f32("full_rows"), fri=i32("full_row_indices"))
    topX = 0 if args["full_rows"] is None else int(args["full_rows"].shape[1])
    if mul_init is not None:
        mul_init = np.ascontiguousarray(np.broadcast_to(np.asarray(mul_init, dtype=np.float32), (batch, N)))
    return bits, K, N, batch, args, topX, vec, mul_init


def forward_f64(layer, vec, mul_init=None):
    """fp64 truth of the whole path (LUT GEMV + CSR + dense rows), accumulating onto mul_init."""
    bits, K, N, batch, a, topX, vec, mul_init = _prep(layer, vec, mul_init)
    out = np.empty((batch, N), dtype=np.float64)
    lib().sq_forward_f64(ctypes.c_int(bits), ctypes.c_int(K), ctypes.c_int(N), ctypes.c_int(batch),
                         _p(a["qweight"]), _p(a["lut"]), _p(a["rows"]), _p(a["cols"]), _p(a["vals"]),
                         _p(a["full_rows"]), _p(a["fri"]), ctypes.c_int(topX),
                         _p(vec), _p(mul_init), _p(out))
    return out


def forward_f32_blocked(layer, vec, mul_init=None):
    """fp32 emulation of one legal execution order of the reference kernels."""
    bits, K, N, batch, a, topX, vec, mul_init = _prep(layer, vec, mul_init)
    mul = np.zeros((batch, N), dtype=np.float32) if mul_init is None else mul_init.copy()
    lib().sq_forward_f32_blocked(ctypes.c_int(bits), ctypes.c_int(K), ctypes.c_int(N), ctypes.c_int(batch),
                                 _p(a["qweight"]), _p(a["lut"]), _p(a["rows"]), _p(a["cols"]), _p(a["vals"]),
                                 _p(a["full_rows"]), _p(a["fri"]), ctypes.c_int(topX),
                                 _p(vec), _p(mul))
    return mul


def dequant_f32(qweight, lut, bits):
    """W[k][c] = LUT[c][idx(k,c)], fp32 [K, N]."""
    qweight = np.ascontiguousarray(qweight, dtype=np.int32)
    lut = np.ascontiguousarray(lut, dtype=np.float32)
    rows, N = qweight.shape
    K = rows * 32 // bits
    W = np.empty((K, N), dtype=np.float32)
    lib().sq_dequant_f32(ctypes.c_int(bits), _p(qweight), _p(lut), ctypes.c_int(K), ctypes.c_int(N), _p(W))
    return W


# ------------------------------------------------------------------------------------------------
# Independent numpy restatement of the format (cross-check for the C code; small/medium sizes)
# ------------------------------------------------------------------------------------------------
def unpack_np(qweight, bits):
    """Vectorised numpy decode following the kernel's expressions
    (squeezellm/quant_cuda_kernel.cu:863-877 for 4-bit, :776-825 for 3-bit)."""
    q = np.ascontiguousarray(qweight).view(np.uint32)
    rows, N = q.shape
    if bits == 4:
        sh = (4 * np.arange(8, dtype=np.uint32))[None, :, None]
        idx = (q[:, None, :] >> sh) & np.uint32(0xF)          # [rows, 8, N]
        return idx.reshape(rows * 8, N).astype(np.uint8)
    assert bits == 3 and rows % 3 == 0
    g = rows // 3
    w0, w1, w2 = q[0::3], q[1::3], q[2::3]                     # each [g, N]
    out = np.empty((g, 32, N), dtype=np.uint32)
    s10 = (3 * np.arange(10, dtype=np.uint32))[None, :, None]
    out[:, 0:10] = (w0[:, None, :] >> s10) & 7
    out[:, 10] = (w0 >> 30) | ((w1 << 2) & 4)
    t = w1 >> 1
    out[:, 11:21] = (t[:, None, :] >> s10) & 7
    out[:, 21] = (t >> 30) | ((w2 << 1) & 6)
    t = w2 >> 2
    out[:, 22:32] = (t[:, None, :] >> s10) & 7
    return out.reshape(g * 32, N).astype(np.uint8)


def pack_np(idx, bits):
    """Vectorised numpy encode following squeezellm/quant.py:177-203."""
    idx = np.asarray(idx).astype(np.uint32)
    K, N = idx.shape
    if bits == 4:
        v = idx.reshape(K // 8, 8, N)
        sh = (4 * np.arange(8, dtype=np.uint32))[None, :, None]
        return np.bitwise_or.reduce(v << sh, axis=1).astype(np.uint32).view(np.int32)
    v = idx.reshape(K // 32, 32, N)
    s10 = (3 * np.arange(10, dtype=np.uint32))[None, :, None]
    w0 = np.bitwise_or.reduce(v[:, 0:10] << s10, axis=1) | (v[:, 10] << 30)
    w1 = ((v[:, 10] >> 2) & 1) | np.bitwise_or.reduce(v[:, 11:21] << (s10 + 1), axis=1) | (v[:, 21] << 31)
    w2 = ((v[:, 21] >> 1) & 3) | np.bitwise_or.reduce(v[:, 22:32] << (s10 + 2), axis=1)
    q = np.empty((K // 32 * 3, N), dtype=np.uint32)
    q[0::3], q[1::3], q[2::3] = w0, w1, w2
    return q.view(np.int32)


# ------------------------------------------------------------------------------------------------
# Synthetic layers in the reference's buffer format
# ------------------------------------------------------------------------------------------------
def make_layer(bits, infeatures, outfeatures, sparsity=0.0, topX=0, seed=0, skew=False,
               nonzero_full_rows=False, bias=False):
    """Random QuantLinearLUT buffers: qweight int32 [K/32*bits, N] (every bit pattern is a valid
    packed weight), lookup_table fp32 [N, 2^bits] (sorted per channel, ~N(0, 0.02)), CSR
    rows/cols/vals with nnz = round(sparsity*K*N) (row = output channel, sorted cols inside a row),
    full_rows fp32 [K, topX] + full_row_indices int32 [topX] (zeros unless nonzero_full_rows, as
    when a checkpoint lacks them - llama.py:182 strict=False)."""
    rng = np.random.default_rng(seed)
    K, N = infeatures, outfeatures
    assert K % 32 == 0
    rows_q = K // 32 * bits
    layer = dict(bits=bits, infeatures=K, outfeatures=N)
    layer["qweight"] = rng.integers(-2**31, 2**31, size=(rows_q, N), dtype=np.int64).astype(np.int32)
    layer["lookup_table"] = np.sort(rng.standard_normal((N, 2**bits)).astype(np.float32) * 0.02, axis=1)
    layer["bias"] = (rng.standard_normal(N).astype(np.float32) * 0.1) if bias else None
    nnz = int(round(sparsity * K * N))
    if nnz > 0:
        if skew:   # half of the non-zeros land in 1 % of the output channels
            hot = rng.choice(N, size=max(1, N // 100), replace=False)
            r_hot = rng.choice(hot, size=nnz // 2)
            r_rest = rng.integers(0, N, size=nnz - nnz // 2)
            r = np.concatenate([r_hot, r_rest])
        else:
            r = rng.integers(0, N, size=nnz)
        counts = np.bincount(r, minlength=N)
        counts = np.minimum(counts, K)
        nnz = int(counts.sum())
        rows = np.zeros(N + 1, dtype=np.int32)
        rows[1:] = np.cumsum(counts)
        cols = np.empty(nnz, dtype=np.int32)
        for c in np.nonzero(counts)[0]:
            n = counts[c]
            cols[rows[c]:rows[c + 1]] = np.sort(_distinct(rng, K, n) if n <= 64
                                                else rng.permutation(K)[:n])
        layer["rows"], layer["cols"] = rows, cols
        layer["vals"] = (rng.standard_normal(nnz) * 0.1).astype(np.float32)
    else:
        layer["rows"] = layer["cols"] = layer["vals"] = None
    if topX > 0:
        if nonzero_full_rows:
            layer["full_rows"] = (rng.standard_normal((K, topX)) * 0.05).astype(np.float32)
            layer["full_row_indices"] = rng.choice(N, size=topX, replace=False).astype(np.int32)
        else:
            layer["full_rows"] = np.zeros((K, topX), dtype=np.float32)
            layer["full_row_indices"] = np.zeros(topX, dtype=np.int32)
    else:
        layer["full_rows"] = layer["full_row_indices"] = None
    return layer


def _distinct(rng, K, n):
    """n distinct ints in [0, K) for small n (rejection)."""
    s = set()
    while len(s) < n:
        s.update(rng.integers(0, K, size=n - len(s)).tolist())
    return np.fromiter(s, dtype=np.int64, count=n)


def make_vec(infeatures, batch=1, seed=0):
    """x as the model produces it: fp16 activations, cast to fp32 (squeezellm/quant.py:223,267)."""
    rng = np.random.default_rng(seed + 12345)
    x = rng.standard_normal((batch, infeatures)).astype(np.float16)
    return x.astype(np.float32)


# ------------------------------------------------------------------------------------------------
# CPU baseline: dequant-to-fp16 + torch.matmul  (BASELINE.json configs[0]; restatement)
# ------------------------------------------------------------------------------------------------
def cpu_dequant_matmul(layer, vec, predequantized=None, compute_dtype="float16"):
    """y = x_fp16 @ W_fp16 (+ CSR + dense rows in fp32), on host cores with torch.

    predequantized: pass a torch fp16 W[K, N] to time the matmul alone ((ii) in BASELINE.md section 4).
    Returns (y float32 [batch, N], W) so callers can reuse W."""
    import torch
    K, N = layer["infeatures"], layer["outfeatures"]
    dt = getattr(torch, compute_dtype)
    if predequantized is None:
        W = torch.from_numpy(dequant_f32(layer["qweight"], layer["lookup_table"], layer["bits"])).to(dt)
    else:
        W = predequantized
    x = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float32).reshape(-1, K))
    y = torch.matmul(x.to(dt), W).float()
    if layer.get("rows") is not None:
        crow = torch.from_numpy(layer["rows"].astype(np.int64))
        col = torch.from_numpy(layer["cols"].astype(np.int64))
        S = torch.sparse_csr_tensor(crow, col, torch.from_numpy(layer["vals"]), size=(N, K))
        y = y + (S @ x.t()).t()
    if layer.get("full_rows") is not None:
        part = x @ torch.from_numpy(layer["full_rows"])          # [batch, topX]
        y.index_add_(1, torch.from_numpy(layer["full_row_indices"].astype(np.int64)), part)
    if layer.get("bias") is not None:
        y = y + torch.from_numpy(layer["bias"])
    return y.numpy(), W
