"""Host restatement of the shared-memory addressing of the tensor-map row passes (csrc/kernels2d_tmap.cuh, tmap_inst.cu):
a 16 x 256 slab lands as 16 boxes of 16 rows x 128 bytes with CU_TENSOR_MAP_SWIZZLE_128B (the 16-byte chunk index XORed
with row & 7).  Checks that the kernel's offset function is that swizzle, that it is a bijection over the slab, and that
the two butterfly passes touch 32 distinct banks per shared-memory wavefront (which is the whole point of the layout)."""
BOX_BYTES = 16 * 128


def tmap_off(r, e):                      # tmap_inst.cu: tmap_off
    return r * 128 + ((((e >> 1) ^ (r & 7)) << 4) | ((e & 1) << 3))


def tmap_pos(l, col):                    # tmap_inst.cu: tmap_pos
    return (col >> 4) * BOX_BYTES + tmap_off(l, col & 15)


def swizzle128(byte_in_box):             # the hardware pattern: address bits 4-6 ^= bits 7-9
    return byte_in_box ^ (((byte_in_box >> 7) & 7) << 4)


def banks(addr, nbytes):
    return {(addr + b) // 4 % 32 for b in range(0, nbytes, 4)}


def test_offset_function_is_the_128b_swizzle_and_a_bijection():
    seen = set()
    for r in range(16):
        for col in range(256):
            linear_in_box = r * 128 + (col & 15) * 8            # what an unswizzled box would hold
            assert tmap_off(r, col & 15) == swizzle128(linear_in_box)
            p = tmap_pos(r, col)
            assert p % 8 == 0 and 0 <= p < 16 * BOX_BYTES
            seen.add(p)
    assert len(seen) == 16 * 256


def test_stride16_pass_is_conflict_free():
    """Thread (row, e) reads element e + 16 k: box k at the same offset.  64-bit accesses are served per half-warp:
    16 lanes = the 16 values of e in one row."""
    for k in range(16):
        for row in range(16):
            used = []
            for e in range(16):
                used.extend(banks(k * BOX_BYTES + tmap_off(row, e), 8))
            assert len(used) == 32 and len(set(used)) == 32


def test_contiguous_pass_is_conflict_free():
    """Thread (box, row) with the row fastest across lanes reads chunk c of its 128-byte row at (c ^ (row & 7)) << 4 with
    128-bit accesses, served per quarter-warp: 8 lanes = 8 consecutive rows."""
    for box in range(16):
        for c in range(8):
            for r0 in (0, 8):
                used = []
                for row in range(r0, r0 + 8):
                    used.extend(banks(box * BOX_BYTES + row * 128 + ((c ^ (row & 7)) << 4), 16))
                assert len(used) == 32 and len(set(used)) == 32
    # and the chunk a thread reads as "c" holds elements 2c, 2c+1 of its row
    for row in range(16):
        for c in range(8):
            assert row * 128 + ((c ^ (row & 7)) << 4) == tmap_off(row, 2 * c)
            assert tmap_off(row, 2 * c + 1) == tmap_off(row, 2 * c) + 8


def test_dense_pitch_would_conflict():
    """The reason for the tensor map: with dense 2048-byte rows the contiguous pass puts the lanes of a quarter-warp
    (thread stride 128 bytes) on the same four banks."""
    used = []
    for t in range(8):
        used.extend(banks(t * 128, 16))
    assert len(set(used)) == 4
