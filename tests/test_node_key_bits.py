"""CPU restatement of the bit tricks in csrc/bvh.cuh's node step (child_slab / leaf_bytes): the multiply that spreads
the non-internal mask to one 0x7f byte per slot, and the PRMT selector that builds (leaf byte << 24 | slot).  The GPU
parity tests (BVH == linear scan) cover the kernel; this pins the constants where no GPU is needed."""


def leaf_bytes(mask):
    lo = ((((mask & 0xF) * 0x00204081) & 0xFFFFFFFF) & 0x01010101) * 0x7F & 0xFFFFFFFF
    hi = (((((mask >> 4) & 0xF) * 0x00204081) & 0xFFFFFFFF) & 0x01010101) * 0x7F & 0xFFFFFFFF
    return lo, hi


def prmt(a, b, sel):
    """PTX prmt.b32 (default mode): result byte i = byte (sel nibble i & 7) of {b, a}; nibble bit 3 replicates its sign."""
    src = [(a >> (8 * k)) & 0xFF for k in range(4)] + [(b >> (8 * k)) & 0xFF for k in range(4)]
    out = 0
    for i in range(4):
        nib = (sel >> (4 * i)) & 0xF
        byte = src[nib & 7]
        if nib & 8:
            byte = 0xFF if byte & 0x80 else 0x00
        out |= byte << (8 * i)
    return out


def test_every_internal_mask_gives_the_expected_key_bits():
    sl = (0x03020100, 0x07060504)
    for imask in range(256):
        lb = leaf_bytes(~imask & 0xFF)
        for s in range(8):
            j = s & 3
            sel = (j << 12) | (0xC << 8) | (0xC << 4) | (4 + j)
            x = prmt(lb[s >> 2], sl[s >> 2], sel)
            internal = (imask >> s) & 1
            assert x == ((0 if internal else 0x7F000000) | s)
            # merged with any non-negative float's bits the key keeps the slot in its low three bits and a leaf's key
            # is above every finite entry distance (< 2^127)
            for tn_bits in (0x00000000, 0x3F800007, 0x7149F2CA):
                key = (tn_bits & ~7 & 0xFFFFFFFF) | x
                assert key & 7 == s and key < 0x80000000
                assert (key >= 0x7F000000) == (not internal)


def test_plane_to_float_selectors():
    """plane_to_float<H>: half-word H of w dropped into the mantissa of 0x43000000 (128 + q / 256 for a 15-bit q)."""
    import struct
    for w in (0x00000000, 0x7FFF0001, 0x12345678, 0x7FFF7FFF):
        for h, sel in ((0, 0x7104), (1, 0x7324)):
            bits = prmt(w, 0x43000000, sel)
            q = (w >> (16 * h)) & 0xFFFF
            assert bits == 0x43000000 | (q << 8)
            assert struct.unpack("<f", struct.pack("<I", bits))[0] == 128.0 + q / 256.0
