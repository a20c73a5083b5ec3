"""Host-side mirror of the reference's RNG (sboxgates.c:246-268): xorshift1024*.

The search functions consume this generator in lock-step with the reference (256 draws on entry to
search_5lut, lut.c:125-135; 512 after phase 1 of search_7lut, lut.c:362-378; one more on success
when the solved inner LUT has don't-care bits, lut.c:104-106), so the caller's later shuffles stay
identical to the reference's.
"""
import struct

MASK64 = (1 << 64) - 1


class Xorshift1024:
    MULT = 1181783497276652981

    def __init__(self, seed):
        """seed: 128 bytes (what the reference reads from /dev/urandom) or 16 ints."""
        if isinstance(seed, (bytes, bytearray)):
            if len(seed) != 128:
                raise ValueError("seed must be 128 bytes")
            self.s = list(struct.unpack("<16Q", bytes(seed)))
        else:
            self.s = [int(x) & MASK64 for x in seed]
            if len(self.s) != 16:
                raise ValueError("seed must have 16 words")
        self.p = 0
        self.draws = 0

    @classmethod
    def from_state(cls, words, p):
        r = cls(list(words))
        r.p = int(p) & 15
        return r

    def next(self):
        s0 = self.s[self.p]
        self.p = (self.p + 1) & 15
        s1 = self.s[self.p]
        s1 ^= (s1 << 31) & MASK64
        self.s[self.p] = s1 ^ s0 ^ (s1 >> 11) ^ (s0 >> 30)
        self.draws += 1
        return (self.s[self.p] * self.MULT) & MASK64

    __call__ = next

    def copy(self):
        """An independent generator in the same state (used to look ahead without consuming)."""
        r = Xorshift1024(list(self.s))
        r.p = self.p
        r.draws = self.draws
        return r
