"""CPU emulation of csrc/prefill_attn.cu's data movement: the XOR-swizzled shared-memory tiles, the per-lane ldmatrix
addresses, the mma.sync m16n8k16 fragment layouts, the S→P register re-use, the causal mask indices and the
double-buffered tile loop — lane by lane, following the kernel statement by statement — checked against the oracle's
flash attention.  It cannot prove the PTX semantics (those are restated here from the PTX ISA's fragment figures; the
A/B/C fragment maps used by mma_16816() below were cross-checked against CUTLASS' own statement of them,
cute/atom/mma_traits_sm80.hpp MMA_Traits<SM80_16x8x16_F32BF16BF16F32_TN>: ALayout/BLayout/SM80_16x8_Row), but it pins
every index formula of the kernel without a GPU.  The GPU run of the real kernel is tests/test_staged_gpu.py.
"""
import math

import numpy as np
import pytest
import torch

from helpers import assert_close_bf16, orc

BM = BN = 64
WARPS = 4


def bf16(x: np.ndarray) -> np.ndarray:
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16).float().numpy()


class Smem:
    """[rows][HD] bf16 tile with 16-byte chunks XOR-swizzled by (row & 7) — sw_off() of the kernel, in elements."""

    def __init__(self, rows, hd):
        self.hd = hd
        self.data = np.full(rows * hd, np.nan, dtype=np.float32)

    def off(self, r, c):
        return r * self.hd + ((c ^ (r & 7)) << 3)

    def store_chunk(self, r, c, vals8):
        o = self.off(r, c)
        self.data[o:o + 8] = vals8

    def chunk_at(self, elem_off):
        return self.data[elem_off:elem_off + 8]


def ldsm_x4(smem: Smem, lane_addr, trans=False):
    """ldmatrix.sync.aligned.m8n8.x4[.trans].b16: lane_addr[l] = element offset of the 16-byte row lane l supplies;
    matrix i is made of the rows supplied by lanes 8i … 8i+7.  Returns regs[lane][i] = (lo, hi) pair."""
    regs = np.zeros((32, 4, 2), dtype=np.float32)
    for i in range(4):
        mat = np.stack([smem.chunk_at(lane_addr[8 * i + r]) for r in range(8)])  # [row][col]
        for t in range(32):
            if not trans:
                regs[t, i] = mat[t // 4, 2 * (t % 4): 2 * (t % 4) + 2]
            else:
                regs[t, i] = (mat[2 * (t % 4), t // 4], mat[2 * (t % 4) + 1, t // 4])
    return regs


def mma_16816(d, a, b0, b1):
    """d[lane][4] += A·B with the PTX fragment layouts: A a0:(g,2t..) a1:(g+8,2t..) a2:(g,2t+8..) a3:(g+8,2t+8..);
    B b0:(k=2t..,n=g) b1:(k=2t+8..,n=g); C/D c0,c1:(g,2t..2t+1) c2,c3:(g+8,2t..2t+1)."""
    A = np.zeros((16, 16), dtype=np.float32)
    B = np.zeros((16, 8), dtype=np.float32)
    for t in range(32):
        g, tq = t // 4, t % 4
        A[g, 2 * tq:2 * tq + 2] = a[t, 0]
        A[g + 8, 2 * tq:2 * tq + 2] = a[t, 1]
        A[g, 2 * tq + 8:2 * tq + 10] = a[t, 2]
        A[g + 8, 2 * tq + 8:2 * tq + 10] = a[t, 3]
        B[2 * tq:2 * tq + 2, g] = b0[t]
        B[2 * tq + 8:2 * tq + 10, g] = b1[t]
    C = A.astype(np.float64) @ B.astype(np.float64)
    for t in range(32):
        g, tq = t // 4, t % 4
        d[t, 0] += C[g, 2 * tq]
        d[t, 1] += C[g, 2 * tq + 1]
        d[t, 2] += C[g + 8, 2 * tq]
        d[t, 3] += C[g + 8, 2 * tq + 1]


def emulate_block(q, k, v, p0, q0, hd, scale):
    """One CTA of attn_prefill_mma_kernel for one head: q [S, hd], k/v [nkeys_total, hd] (bf16 values as fp32)."""
    S = q.shape[0]
    CH, KS, ND = hd // 8, hd // 16, hd // 8
    rows_valid = min(BM, S - q0)
    nkeys = p0 + q0 + rows_valid
    T = (nkeys + BN - 1) // BN
    Qs = Smem(BM, hd)
    Ks = [Smem(BN, hd), Smem(BN, hd)]
    Vs = [Smem(BN, hd), Smem(BN, hd)]
    for i in range(BM * CH):
        r, c = i // CH, i % CH
        Qs.store_chunk(r, c, q[q0 + r, c * 8:c * 8 + 8] if r < rows_valid else np.zeros(8, np.float32))

    def load_kv(t, buf):
        for i in range(BN * CH):
            r, c = i // CH, i % CH
            key = t * BN + r
            if key < nkeys:
                Ks[buf].store_chunk(r, c, k[key, c * 8:c * 8 + 8])
                Vs[buf].store_chunk(r, c, v[key, c * 8:c * 8 + 8])
            else:
                Ks[buf].store_chunk(r, c, np.zeros(8, np.float32))
                Vs[buf].store_chunk(r, c, np.zeros(8, np.float32))

    out = np.zeros((BM, hd), dtype=np.float32)
    lanes = np.arange(32)
    load_kv(T - 1, 0)
    state = []
    for warp in range(WARPS):
        state.append(dict(o=np.zeros((ND, 32, 4), np.float32), m=np.full((32, 2), -np.inf, np.float32),
                          l=np.zeros((32, 2), np.float32), qf=None))
    for it in range(T):
        t, buf = T - 1 - it, it & 1
        if it + 1 < T:
            load_kv(t - 1, buf ^ 1)
        for warp in range(WARPS):
            st = state[warp]
            g, tq = lanes // 4, lanes % 4
            rpos = [p0 + q0 + np.minimum(warp * 16 + g, rows_valid - 1),
                    p0 + q0 + np.minimum(warp * 16 + g + 8, rows_valid - 1)]
            if it == 0:
                st["qf"] = [ldsm_x4(Qs, [Qs.off(warp * 16 + (l & 15), 2 * kk + (l >> 4)) for l in range(32)])
                            for kk in range(KS)]
            s = np.zeros((8, 32, 4), np.float32)
            for kk in range(KS):
                for n2 in range(4):
                    kb = ldsm_x4(Ks[buf], [Ks[buf].off(n2 * 16 + (l & 7) + ((l >> 4) << 3), 2 * kk + ((l >> 3) & 1))
                                           for l in range(32)])
                    mma_16816(s[2 * n2], st["qf"][kk], kb[:, 0], kb[:, 1])
                    mma_16816(s[2 * n2 + 1], st["qf"][kk], kb[:, 2], kb[:, 3])
            kbase = t * BN
            if kbase + BN - 1 > p0 + q0:
                for n in range(8):
                    j = kbase + n * 8 + 2 * tq
                    s[n, j > rpos[0], 0] = -np.inf
                    s[n, j + 1 > rpos[0], 1] = -np.inf
                    s[n, j > rpos[1], 2] = -np.inf
                    s[n, j + 1 > rpos[1], 3] = -np.inf
            pa = np.zeros((4, 32, 4, 2), np.float32)
            for rr in range(2):
                mx = np.max(np.maximum(s[:, :, 2 * rr], s[:, :, 2 * rr + 1]), axis=0)      # per lane
                mx = mx.reshape(8, 4).max(axis=1).repeat(4)                               # quad shuffle-xor 1, 2
                m_new = np.maximum(st["m"][:, rr], mx)
                with np.errstate(invalid="ignore"):
                    alpha = np.where(np.isneginf(st["m"][:, rr]), 0.0,
                                     np.exp2((st["m"][:, rr] - m_new) * scale)).astype(np.float32)
                m_scaled = np.where(np.isneginf(m_new), 0.0, m_new * scale).astype(np.float32)
                ssum = np.zeros(32, np.float32)
                for n in range(8):
                    p0v = np.exp2(s[n, :, 2 * rr] * scale - m_scaled).astype(np.float32)
                    p1v = np.exp2(s[n, :, 2 * rr + 1] * scale - m_scaled).astype(np.float32)
                    ssum += p0v + p1v
                    pa[n >> 1, :, (n & 1) * 2 + rr, 0] = bf16(p0v)
                    pa[n >> 1, :, (n & 1) * 2 + rr, 1] = bf16(p1v)
                st["l"][:, rr] = st["l"][:, rr] * alpha + ssum
                st["m"][:, rr] = m_new
                st["o"][:, :, 2 * rr] *= alpha
                st["o"][:, :, 2 * rr + 1] *= alpha
            for k2 in range(4):
                for d2 in range(ND // 2):
                    vb = ldsm_x4(Vs[buf], [Vs[buf].off(k2 * 16 + (l & 7) + (((l >> 3) & 1) << 3), 2 * d2 + (l >> 4))
                                           for l in range(32)], trans=True)
                    mma_16816(st["o"][2 * d2], pa[k2], vb[:, 0], vb[:, 1])
                    mma_16816(st["o"][2 * d2 + 1], pa[k2], vb[:, 2], vb[:, 3])
    for warp in range(WARPS):
        st = state[warp]
        for rr in range(2):
            l = st["l"][:, rr].reshape(8, 4).sum(axis=1).repeat(4)
            inv = np.where(l > 0, 1.0 / np.where(l > 0, l, 1.0), 0.0).astype(np.float32)
            for lane in range(32):
                g, tq = lane // 4, lane % 4
                r = warp * 16 + g + 8 * rr
                if r < rows_valid:
                    for n in range(ND):
                        out[r, n * 8 + 2 * tq] = st["o"][n, lane, 2 * rr] * inv[lane]
                        out[r, n * 8 + 2 * tq + 1] = st["o"][n, lane, 2 * rr + 1] * inv[lane]
    return bf16(out[:rows_valid])


@pytest.mark.parametrize("hd,S,p0,Hq,Hkv", [(64, 70, 0, 2, 1), (64, 64, 37, 1, 1), (128, 33, 100, 2, 2), (128, 130, 0, 2, 1)])
def test_kernel_index_math_against_oracle(hd, S, p0, Hq, Hkv):
    gen = torch.Generator().manual_seed(hd + S + p0)
    total = p0 + S
    q = torch.randn(1, S, Hq, hd, generator=gen).to(torch.bfloat16)
    k = torch.randn(1, total, Hkv, hd, generator=gen).to(torch.bfloat16)
    v = torch.randn(1, total, Hkv, hd, generator=gen).to(torch.bfloat16)
    scale = np.float32(np.float32(1.0 / math.sqrt(hd)) * np.float32(1.4426950408889634))
    # oracle: each query row r sees keys 0 … p0 + r  (top-left aligned causal mask once the prefix rows are prepended
    # as "virtual" query rows — run the oracle on the full square problem and keep the last S rows)
    q_full = torch.cat([torch.zeros(1, p0, Hq, hd, dtype=torch.bfloat16), q], dim=1)
    want = orc.flash_attention(q_full, k, v, True, "bf16")[0, p0:]
    got = np.zeros((S, Hq, hd), dtype=np.float32)
    for h in range(Hq):
        kvh = h // (Hq // Hkv)
        for q0 in range(0, S, BM):
            blk = emulate_block(q[0, :, h].float().numpy(), k[0, :, kvh].float().numpy(), v[0, :, kvh].float().numpy(),
                                p0, q0, hd, scale)
            got[q0:q0 + blk.shape[0], h] = blk
    # tile size and visiting order equal the oracle's only by accident: allow the same 4 ulp as the GPU attention tests
    assert_close_bf16(torch.from_numpy(got), want, 4, f"emulated prefill attention hd={hd} S={S} p0={p0}", atol=2e-3)
