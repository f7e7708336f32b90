"""Would a tensor-core Gram tile hold the parity tolerance for the wide Euclidean cost (configs[2]: 1024 x 4950 pair distances)?

Emulates on the CPU what a tcgen05 kernel would compute -- D2_ij = |x_i|^2 + |x_j|^2 - 2 x_i.x_j with the dot products from
split-precision tensor-core passes accumulated in float32 -- and compares loss and dL/dz of the sketch-map cost with the
float64 oracle (tolerance 1e-5 relative, BASELINE.json).  Formats: TF32 (10-bit mantissa) and BF16 (7-bit), with 1, 2 (hi*hi +
hi*lo + lo*hi) and 3-term splits, with and without centring the rows first; for reference the float32 Gram form the
REFERENCE itself uses (encodermap/misc/distances.py:211-233) and the float32 difference form libemk uses.

    python tools/experiments/gram_split_precision.py  -> profiles/r02_gram_split_precision.txt
"""
import math
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from oracle import em_oracle as O  # noqa: E402

SIG = (4.5, 12, 6, 1, 2, 6)


def trunc_mantissa(x32: np.ndarray, bits: int) -> np.ndarray:
    """round-to-nearest-even float32 -> float with `bits` explicit mantissa bits (TF32: 10, BF16: 7), kept in float32"""
    u = x32.view(np.uint32).astype(np.uint64)
    drop = 23 - bits
    half = np.uint64(1 << (drop - 1))
    lsb = (u >> np.uint64(drop)) & np.uint64(1)
    u = (u + half - np.uint64(1) + lsb) >> np.uint64(drop) << np.uint64(drop)
    return u.astype(np.uint32).view(np.float32)


def split(x32, bits, terms):
    parts, rest = [], x32.copy()
    for _ in range(terms):
        p = trunc_mantissa(rest, bits)
        parts.append(p)
        rest = (rest - p).astype(np.float32)
    return parts


def gram_split(x32, bits, terms):
    """x x^T from tensor-core passes: products of `bits`-mantissa operands are exact, accumulation in float32 per pass
    (emulated with a float32 matmul of the rounded operands: the operand rounding is what matters here)"""
    parts = split(x32, bits, terms)
    g = np.zeros((x32.shape[0], x32.shape[0]), np.float32)
    for a in range(terms):
        for b in range(terms):
            if a + b < terms:     # hi*hi, hi*lo, lo*hi, (hi*lo2, lo*lo, lo2*hi) ...
                g += parts[a] @ parts[b].T
    return g


def cost_from_d2(d2h, low, sig):
    """loss and dL/dz in float64 given the high-d squared distances (everything else exact: isolates the Gram error)"""
    z = torch.from_numpy(low).double().requires_grad_(True)
    dh = torch.sqrt(torch.clamp_min(torch.from_numpy(d2h.astype(np.float64)), 0.0))
    dl = O.pairwise_dist(z)[0]
    loss = torch.mean((O.sigmoid(*sig[:3])(dh) - O.sigmoid(*sig[3:])(dl)) ** 2)
    loss.backward()
    return loss.item(), z.grad.numpy()


def main(out):
    rng = np.random.default_rng(1024 + 4950)
    n, d = 1024, 4950
    centres = rng.uniform(0.4, 8.0, size=(8, d))
    x = (centres[rng.integers(0, 8, n)] + rng.normal(scale=4.5 / math.sqrt(2 * d), size=(n, d))).astype(np.float32)
    low = (rng.normal(size=(n, 2)) * 1.5).astype(np.float32)
    x64 = x.astype(np.float64)
    sq = (x64 * x64).sum(1)
    d2_exact = np.maximum(sq[:, None] + sq[None, :] - 2 * x64 @ x64.T, 0)
    l0, g0 = cost_from_d2(d2_exact, low, SIG)
    rows = []

    def report(name, d2):
        l, g = cost_from_d2(d2, low, SIG)
        near = d2_exact < 4 * SIG[0] ** 2                  # pairs inside the sigmoid's active range
        np.fill_diagonal(near, False)
        rel_d2 = np.abs(d2 - d2_exact)[near] / d2_exact[near]
        rows.append((name, abs(l - l0) / l0, np.linalg.norm(g - g0) / np.linalg.norm(g0), float(np.median(rel_d2)), float(rel_d2.max())))

    # float32 difference form (libemk): sum_k (x_ik - x_jk)^2 in float32
    d2 = np.zeros((n, n), np.float32)
    for i0 in range(0, n, 64):
        diff = x[i0:i0 + 64, None, :] - x[None, :, :]
        d2[i0:i0 + 64] = np.einsum("ijk,ijk->ij", diff, diff, dtype=np.float32)
    report("float32 difference form (libemk FFMA kernel)", d2)
    for centre in (False, True):
        xc = (x - x.mean(0, keepdims=True).astype(np.float32)) if centre else x
        tag = " + centred rows" if centre else ""
        sq32 = (xc.astype(np.float64) ** 2).sum(1)        # norms in float64: only the dot products come from tensor cores
        g32 = xc @ xc.T
        report("float32 Gram form (the reference's own)" + tag, sq32[:, None] + sq32[None, :] - 2 * g32.astype(np.float64))
        for fmt, bits in (("TF32", 10), ("BF16", 7)):
            for terms in (1, 2, 3):
                g = gram_split(xc, bits, terms)
                report(f"{fmt} x{terms} split Gram, fp32 accumulate" + tag, sq32[:, None] + sq32[None, :] - 2 * g.astype(np.float64))
    lines = ["# tools/experiments/gram_split_precision.py: configs[2] shape 1024 x 4950, clustered pair distances (8 clusters, sigma_h = 4.5)",
             f"# float64 loss {l0:.6e}; tolerance on loss and gradient: 1e-5 relative (BASELINE.json)",
             f"{'high-d squared distances from':62s} {'loss rel err':>12s} {'grad rel err':>12s} {'median dD2/D2':>14s} {'max dD2/D2':>12s}  (near pairs)"]
    for name, el, eg, med, mx in rows:
        ok = "ok " if max(el, eg) < 1e-5 else "FAIL"
        lines.append(f"{name:62s} {el:12.2e} {eg:12.2e} {med:14.2e} {mx:12.2e}  {ok}")
    text = "\n".join(lines)
    print(text)
    if out:
        Path(out).write_text(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
