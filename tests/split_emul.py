"""CPU restatement (numpy, exact integer arithmetic) of the tcgen05 split-integer composition GEMM
(analyticmesh_b200/csrc/split.cuh).  Test infrastructure only: the GPU result must equal this bit for bit.

    digits:   X = rint(x * 2^(8 SD - 2 - e)),  2^e > max |x| over the weight row / masked plane column,
              X = sum_t d_t 256^(SD-1-t) with d_t in [-128, 127]  (bytes of (X + 0x80..80) ^ 0x80..80)
    products: D_g = sum_{i+j=g} dA_i . dB_j            (exact, g < SD)
    result:   acc = D_{SD-1};  acc = acc * 2^-8 + D_g  (g = SD-2 .. 0, one rounding per step)
              out = acc * 2^(eA-6) * 2^(eB-6)
"""
import numpy as np


def _exponent(mx):
    """e with mx = f * 2^e, f in [0.5, 1); 0 where mx == 0."""
    _, e = np.frexp(mx)
    return np.where((mx >= np.finfo(np.float64).tiny) & np.isfinite(mx), e, 0).astype(np.int64)


def digits(x, e, sd):
    """x: float64 array, e: broadcastable exponents -> int64 array (sd, *x.shape) of digits in [-128, 127]."""
    X = np.rint(np.ldexp(x, (8 * sd - 2) - e)).astype(np.int64)
    C = np.uint64(int("80" * sd, 16))
    Y = (X.astype(np.uint64) + C) ^ C
    out = np.empty((sd,) + x.shape, dtype=np.int64)
    for t in range(sd):
        out[t] = ((Y >> np.uint64(8 * (sd - 1 - t))) & np.uint64(0xFF)).astype(np.uint8).view(np.int8).astype(np.int64)
    return out


def split_gemm(W, B, sd=7):
    """W (M, K) float64, B (K, N) float64 (already masked) -> (M, N) float64, as the tensor-core path computes it."""
    eA = _exponent(np.abs(W).max(axis=1))
    eB = _exponent(np.abs(B).max(axis=0))
    dA = digits(W, eA[:, None], sd)
    dB = digits(B, eB[None, :], sd)
    D = [sum(dA[i] @ dB[g - i] for i in range(g + 1)) for g in range(sd)]
    assert max(int(np.abs(d).max()) for d in D) < 2 ** 31
    acc = D[sd - 1].astype(np.float64)
    for g in range(sd - 2, -1, -1):
        acc = acc * 2.0 ** -8 + D[g].astype(np.float64)
    return acc * np.ldexp(1.0, eA - 6)[:, None] * np.ldexp(1.0, eB - 6)[None, :]


def compose(info, state_bits, sd=7):
    """Unsigned plane rows (L, 4) of one activation pattern, composed layer by layer with split_gemm
    (same layer / skip semantics as the engine's compose_chunk; reference backend/inc/process.h:18-193)."""
    nodes = info.nodes
    D = len(nodes) - 2
    off = np.concatenate([[0], np.cumsum(nodes[1:-1])]).astype(int)       # off[h-1] = first bit of hidden layer h
    W = [w.astype(np.float64) for w in info.weights]
    b = [v.astype(np.float64) for v in info.biases]
    rows = {1: np.concatenate([W[0], b[0][:, None]], axis=1)}
    bits = {h: state_bits[off[h - 1]:off[h]].astype(np.float64) for h in range(1, D + 1)}
    for h in range(1, D):
        out = split_gemm(W[h], rows[h] * bits[h][:, None], sd)
        out[:, 3] += b[h]
        row = info.arc_table[h - 1]                                         # skips into hidden layer h + 1
        for j in range(int(row[0])):
            src, tm = int(row[1 + 2 * j]), int(row[2 + 2 * j])
            T = info.arc_tm[tm].astype(np.float64)
            if src == 0:
                if T.size == 0:
                    for m in range(min(3, out.shape[0])):
                        out[m, m] += 1.0
                else:
                    out[:, :3] += T
            else:
                masked = rows[src] * bits[src][:, None]
                out = out + (masked if T.size == 0 else split_gemm(T, masked, sd))
        rows[h + 1] = out
    return np.concatenate([rows[h] for h in range(1, D + 1)], axis=0)
