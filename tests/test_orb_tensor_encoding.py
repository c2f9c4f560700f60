"""CPU checks of the arithmetic the ORB tensor-core route relies on (csrc/sift_prep.cu
orb_tc_prep_kernel, csrc/sift_tc.cu): Hamming distance as a squared L2 distance of 0/1 vectors, the
e4m3 byte values used for the operands, and the exact three-piece encoding of popcount / 2 in the
augmentation block.  The kernels themselves are tested on the GPU against the oracle
(tests/test_gpu_matching.py::test_orb_both_kernels_*)."""
import numpy as np

from oracle import c_oracle, synth


def e4m3_decode(b):
    """OCP FP8 E4M3 (bias 7, no infinities, 0x7F / 0xFF = NaN)."""
    s = -1.0 if b & 0x80 else 1.0
    e, m = (b >> 3) & 0xF, b & 7
    if e == 0xF and m == 7:
        return float("nan")
    if e == 0:
        return s * m * 2.0 ** -9
    return s * (1 + m / 8.0) * 2.0 ** (e - 7)


def e4m3_encode_exact(v):
    """The byte whose value is exactly v (the prep kernel only ever encodes representable values)."""
    for b in range(0x7F):
        if e4m3_decode(b) == v:
            return b
    raise AssertionError(f"{v} is not an e4m3 value")


def test_operand_bytes():
    assert e4m3_decode(0x00) == 0.0 and e4m3_decode(0x38) == 1.0      # a descriptor bit
    assert e4m3_decode(0x30) == 0.5                                    # the odd-popcount piece
    assert e4m3_decode(0x7E) == 448.0                                  # padding rows
    assert max(e4m3_decode(b) for b in range(0x7F)) == 448.0          # ... the largest finite value


def test_popcount_half_is_three_exact_pieces():
    for pc in range(257):
        h, m, l = (pc >> 5) << 4, (pc >> 1) & 15, 0.5 * (pc & 1)      # orb_tc_prep_kernel's split
        assert h + m + l == pc / 2
        for piece in (float(h), float(m), l):
            assert e4m3_decode(e4m3_encode_exact(piece)) == piece


def test_hamming_is_squared_l2_of_bit_vectors():
    q, t = synth.orb_pair(64, 96, 77)
    qb = np.unpackbits(q, axis=1).astype(np.float32)
    tb = np.unpackbits(t, axis=1).astype(np.float32)
    # what the tensor cores accumulate: -(q.t) + |q|^2/2 + |t|^2/2, exact in fp32 (all terms are
    # multiples of 1/2 below 2^9)
    acc = -(qb @ tb.T) + qb.sum(1)[:, None] / 2 + tb.sum(1)[None, :] / 2
    ham = np.bitwise_count(q[:, None, :] ^ t[None, :, :]).sum(2)
    assert np.array_equal(2 * acc, ham.astype(np.float32))
    idx, dist = c_oracle.hamming_knn2(q, t)
    assert np.array_equal(dist[:, 0], ham.min(1).astype(np.float32))
    assert np.array_equal(idx[:, 0], ham.argmin(1))                    # ties: lowest train index


def test_padding_threshold_separates_real_from_padding():
    # a real accumulator is Hamming / 2 <= 128; a padding column adds 448 to popcount(q) / 2 >= 0;
    # sift_tc.cu reads anything above ORB_PAD_THRESHOLD = 200 as "no such column"
    assert 256 / 2 < 200 < 448
