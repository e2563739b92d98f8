"""Oracle (test infrastructure): the reference's default food list from first principles.

structs.jl:33,70 — every SnakeGame() draws 50 cells (rand(2:9), rand(2:9)), row before
column, from a fresh Xoshiro(42).  Julia 1.10 seeds Xoshiro(seed::Integer) with the four
little-endian UInt64 words of SHA-256 over the seed's UInt32 limbs (Random/src/Xoshiro.jl,
Random.seed!/hash_seed), generator xoshiro256++, and rand(a:b) over a power-of-two range of
length 8 reduces to the top three bits of one 64-bit draw (SamplerRangeNDL: (x*8)>>64).

Pinned by tests/test_oracle_golden.py against the list AND the post-draw generator state
stored in the reference's trainers/very_long_training1.bson (tests/golden/g1_food_list.json).
"""
import hashlib
import struct

M64 = (1 << 64) - 1


def _rotl(x, k):
    return ((x << k) | (x >> (64 - k))) & M64


class Xoshiro256pp:
    def __init__(self, seed: int):
        limbs = struct.pack("<I", seed & 0xFFFFFFFF)          # 42 -> 2a 00 00 00
        self.s = list(struct.unpack("<4Q", hashlib.sha256(limbs).digest()))

    def next_u64(self) -> int:
        s0, s1, s2, s3 = self.s
        res = (_rotl((s0 + s3) & M64, 23) + s0) & M64
        t = (s1 << 17) & M64
        s2 ^= s0
        s3 ^= s1
        s1 ^= s2
        s0 ^= s3
        s2 ^= t
        s3 = _rotl(s3, 45)
        self.s = [s0, s1, s2, s3]
        return res

    def rand_range8(self, lo: int) -> int:
        """rand(lo:lo+7)"""
        return (self.next_u64() >> 61) + lo


def default_food_list(seed: int = 42, n: int = 50):
    """[(row, col)] 1-based, and the generator state after the 2n draws."""
    g = Xoshiro256pp(seed)
    cells = []
    for _ in range(n):
        r = g.rand_range8(2)
        c = g.rand_range8(2)
        cells.append((r, c))
    return cells, list(g.s)


if __name__ == "__main__":
    cells, st = default_food_list()
    print(cells)
    print(["0x%016x" % w for w in st])
