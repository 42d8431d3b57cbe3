"""CPU check of the arithmetic the tensor-core act kernels use (csrc/zbot_policy_tc.cuh, zbot_policy_tc5.cuh): every FP32 operand
x is split as hi + lo (hi = x rounded to TF32 by adding half an ulp of the kept field to the bit pattern and clearing the 13
dropped bits; lo = x - hi, of which the tensor core reads the upper 19 bits), and a product is lo_a*hi_b + hi_a*lo_b + hi_a*hi_b
accumulated in FP32.  A numpy emulation of the whole actor network (num_obs -> 128 -> 128 -> 128 -> 6, ELU) -- with a PESSIMISTIC
accumulator that truncates toward zero after every group of eight products -- must stay at FP32 round-off from float64, far from
what plain TF32 products give.  (The kernels themselves are held to the same bounds on the GPU: tests/test_gpu_policy.py.)"""
import numpy as np


def tf32_hi(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


def trunc19(x):
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def acc_toward_zero(acc, add64):
    t = acc.astype(np.float64) + add64
    t32 = t.astype(np.float32)
    over = np.abs(t32.astype(np.float64)) > np.abs(t)
    t32[over] = np.nextafter(t32[over], np.float32(0))
    return t32


def matmul_split(a, b, terms):
    ah, bh = tf32_hi(a), tf32_hi(b)
    al, bl = trunc19((a - ah).astype(np.float32)), trunc19((b - bh).astype(np.float32))
    acc = np.zeros((a.shape[0], b.shape[1]), np.float32)
    for k0 in range(0, a.shape[1], 8):
        s = slice(k0, k0 + 8)
        pairs = ((al, bh), (ah, bl), (ah, bh)) if terms == 3 else ((ah, bh),)
        for p, q in pairs:
            acc = acc_toward_zero(acc, p[:, s].astype(np.float64) @ q[s].astype(np.float64))
    return acc


def elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0))).astype(x.dtype)


def test_three_term_tf32_split_products_are_fp32_accurate():
    rng = np.random.default_rng(0)
    n, dims = 1024, [23, 128, 128, 128, 6]
    Ws = [(rng.uniform(-1, 1, (dims[i + 1], dims[i])) / np.sqrt(dims[i])).astype(np.float32) for i in range(4)]
    bs = [rng.uniform(-0.3, 0.3, dims[i + 1]).astype(np.float32) for i in range(4)]
    x = (rng.standard_normal((n, dims[0])) * 1.5).astype(np.float32)

    def forward(mm, dt):
        h = x.astype(dt)
        for i in range(4):
            h = mm(i, h) + bs[i].astype(dt)
            if i < 3:
                h = elu(h).astype(dt)
        return h

    ref64 = forward(lambda i, h: h @ Ws[i].T.astype(np.float64), np.float64)
    ref32 = forward(lambda i, h: (h @ Ws[i].T).astype(np.float32), np.float32)

    def split_mm(terms):
        def mm(i, h):
            if i == 3:                                  # the head runs in plain FP32 in the kernels
                return (h @ Ws[i].T).astype(np.float32)
            k8 = (h.shape[1] + 7) // 8 * 8
            a = np.zeros((n, k8), np.float32)
            a[:, :h.shape[1]] = h
            b = np.zeros((k8, Ws[i].shape[0]), np.float32)
            b[:h.shape[1]] = Ws[i].T
            return matmul_split(a, b, terms)
        return mm

    e_fp32 = np.abs(ref32 - ref64).max()
    e_split = np.abs(forward(split_mm(3), np.float32) - ref64).max()
    e_tf32 = np.abs(forward(split_mm(1), np.float32) - ref64).max()
    assert e_fp32 < 1e-6
    assert e_split < 2e-6 and e_split < 4 * e_fp32 + 2e-6, (e_split, e_fp32)      # the GPU test's bound
    assert e_tf32 > 50 * e_split, (e_tf32, e_split)                                # plain TF32 products are NOT good enough
    # the hi part is TF32 (13 low bits clear) and hi + lo restores the operand exactly
    v = (rng.standard_normal(4096) * 3).astype(np.float32)
    hi = tf32_hi(v)
    assert np.all(hi.view(np.uint32) & np.uint32(0x1FFF) == 0)
    assert np.array_equal((hi + (v - hi).astype(np.float32)).astype(np.float32), v)
