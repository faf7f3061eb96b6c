"""The tensor-core head (csrc/cs_head_mma.cuh) feeds fp32 operands to TF32 MMAs as two terms:
big = x rounded to 10 explicit mantissa bits (integer add of half an ulp, then mask -- what
`split_tf32` does) and small = x - big handed over as is (the tensor core ignores its low 13 mantissa
bits), and accumulates small*big + big*small + big*big.  This is a bit-level emulation of that scheme
in numpy: the reconstructed product must carry fp32-class accuracy (DESIGN.md section 7 quotes
2^-21 per operand), which is what lets the head keep the parity tolerances of the SIMT kernel."""
import numpy as np


def _split(x):
    bits = x.view(np.uint32)
    big = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    small = (x - big).astype(np.float32)                       # exact in fp32
    small_seen = (small.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)   # hardware truncation
    return big, small_seen


def test_three_term_tf32_product_is_fp32_accurate():
    rng = np.random.default_rng(0)
    a = (rng.standard_normal(1 << 16) * np.exp(rng.uniform(-8, 8, 1 << 16))).astype(np.float32)
    b = (rng.standard_normal(1 << 16) * np.exp(rng.uniform(-8, 8, 1 << 16))).astype(np.float32)
    ab, as_ = _split(a)
    bb, bs = _split(b)
    # operand reconstruction error: |x - (big + small_seen)| <= 2^-21 |x|
    for x, big, small in ((a, ab, as_), (b, bb, bs)):
        rel = np.abs(x.astype(np.float64) - (big.astype(np.float64) + small.astype(np.float64))) / np.abs(x)
        assert rel.max() <= 2.0 ** -21, rel.max()
    exact = a.astype(np.float64) * b.astype(np.float64)
    three = (as_.astype(np.float64) * bb + ab.astype(np.float64) * bs + ab.astype(np.float64) * bb)
    rel = np.abs(three - exact) / np.abs(exact)
    # dropped term small*small <= 2^-22, two truncations <= 2^-21 each
    assert rel.max() <= 2.0 ** -19.5, rel.max()
    assert np.sqrt((rel ** 2).mean()) <= 2.0 ** -21.5
    # a single TF32 product (what the kernel would lose without the split) is ~1000x worse
    one = ab.astype(np.float64) * bb
    assert (np.abs(one - exact) / np.abs(exact)).max() > 2.0 ** -12


def test_big_is_round_to_nearest_on_ten_mantissa_bits():
    x = np.array([1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -11 + 2.0 ** -20, 1.0 + 2.0 ** -10, -3.14159274, 1e-30, 6.5e4],
                 dtype=np.float32)
    big, _ = _split(x)
    for v, g in zip(x, big):
        m = np.frexp(g)[0]
        assert (abs(m) * 2 ** 11) % 1 == 0                     # 11 significant bits
        assert abs(float(v) - float(g)) <= abs(float(v)) * 2.0 ** -11
    assert big[1] == np.float32(1.0 + 2.0 ** -10)              # a tie rounds away from zero, like cvt.rna
