"""ORACLE (test infrastructure, not product code): range-ANS entropy coder restated in plain Python integers.

The reference codes its latents with CompressAI's `ans` extension (reference main/model/pnet.py:45-49,69-73 ->
`Cheng2020Anchor.compress` -> `BufferedRansEncoder` / `RansEncoder`), a C++ wrapper around the public-domain
`ryg_rans` 64-bit coder.  CompressAI is NOT in /root/reference and is unpinned (requirement.txt:8, ~1.1.x), so this
file restates the published algorithm:

* rans64 (ryg_rans `rans64.h`): 64-bit state, lower bound L = 2^31, 32-bit renormalisation words, symbols pushed in
  REVERSE order, the stream is the words in the order the decoder reads them, little-endian, length a multiple of 4
  bytes and at least 8 (the final state).
* CompressAI `rans_interface.cpp`: 16-bit probability precision; per symbol a CDF row chosen by `indexes`, an
  `offset` subtracted, values outside [0, max_value) escape through the row's last bin followed by a bypass code
  (4-bit digits: digit count in unary-by-15, then the digits of the folded value, least significant first).
* `pmf_to_quantized_cdf` (CompressAI `_CXX`): round(p * 2^16), renormalise to 2^16, steal from the smallest bin > 1
  for every empty bin.

PARITY UNPINNED: nothing in /root/reference pins a byte of these streams; the round trip (decode(encode(s)) == s) and
the stream length against the ideal code length are what the tests check, and the CUDA/host product path is held to
byte identity with THIS restatement.
"""
import math

RANS64_L = 1 << 31
PRECISION = 16
BYPASS_PRECISION = 4
MAX_BYPASS_VAL = (1 << BYPASS_PRECISION) - 1
_M32 = 0xFFFFFFFF


def pmf_to_quantized_cdf(pmf, precision=PRECISION):
    """List of float32 probabilities (last entry = tail mass) -> list of len(pmf)+1 cumulative frequencies."""
    import numpy as np
    p = np.asarray(pmf, dtype=np.float32) * np.float32(1 << precision)
    # std::round: half away from zero (numpy's round is half-to-even)
    q = np.floor(np.abs(p) + np.float32(0.5)).astype(np.int64)
    cdf = [0] + [int(v) for v in q]
    total = sum(cdf)
    if total == 0:
        raise ValueError("pmf sums to zero")
    cdf = [((1 << precision) * v) // total for v in cdf]
    for i in range(1, len(cdf)):
        cdf[i] += cdf[i - 1]
    cdf[-1] = 1 << precision
    n = len(cdf)
    for i in range(n - 1):
        if cdf[i] == cdf[i + 1]:
            best_freq, best_steal = 1 << 32, -1
            for j in range(n - 1):
                freq = cdf[j + 1] - cdf[j]
                if 1 < freq < best_freq:
                    best_freq, best_steal = freq, j
            if best_steal < 0:
                raise ValueError("no bin to steal from")
            if best_steal < i:
                for j in range(best_steal + 1, i + 1):
                    cdf[j] -= 1
            else:
                for j in range(i + 1, best_steal + 1):
                    cdf[j] += 1
    return cdf


def _rans_symbols(symbols, indexes, cdfs, cdf_lengths, offsets):
    """rans_interface.cpp `encode_with_indexes`: (start, range, bypass) triples in coding order."""
    out = []
    for s, ci in zip(symbols, indexes):
        cdf = cdfs[ci]
        max_value = cdf_lengths[ci] - 2
        value = s - offsets[ci]
        raw = 0
        if value < 0:
            raw = -2 * value - 1
            value = max_value
        elif value >= max_value:
            raw = 2 * (value - max_value)
            value = max_value
        out.append((cdf[value], cdf[value + 1] - cdf[value], False))
        if value == max_value:
            n_bypass = 0
            while (raw >> (n_bypass * BYPASS_PRECISION)) != 0:
                n_bypass += 1
            val = n_bypass
            while val >= MAX_BYPASS_VAL:
                out.append((MAX_BYPASS_VAL, MAX_BYPASS_VAL + 1, True))
                val -= MAX_BYPASS_VAL
            out.append((val, val + 1, True))
            for j in range(n_bypass):
                v = (raw >> (j * BYPASS_PRECISION)) & MAX_BYPASS_VAL
                out.append((v, v + 1, True))
    return out


def encode_with_indexes(symbols, indexes, cdfs, cdf_lengths, offsets):
    """-> bytes.  Same result for CompressAI's RansEncoder and BufferedRansEncoder.flush()."""
    syms = _rans_symbols(symbols, indexes, cdfs, cdf_lengths, offsets)
    x = RANS64_L
    words = []  # in push order; the stream is the reverse
    for start, rng, bypass in reversed(syms):
        if not bypass:
            x_max = ((RANS64_L >> PRECISION) << 32) * rng
            if x >= x_max:
                words.append(x & _M32)
                x >>= 32
            x = ((x // rng) << PRECISION) + (x % rng) + start
        else:
            freq = 1 << (16 - BYPASS_PRECISION)
            x_max = ((RANS64_L >> 16) << 32) * freq
            if x >= x_max:
                words.append(x & _M32)
                x >>= 32
            x = (x << BYPASS_PRECISION) | start
    words.append((x >> 32) & _M32)
    words.append(x & _M32)
    words.reverse()
    return b"".join(w.to_bytes(4, "little") for w in words)


class Decoder:
    """rans_interface.cpp `RansDecoder` (set_stream / decode_stream)."""

    def __init__(self, data):
        self.w = [int.from_bytes(data[i:i + 4], "little") for i in range(0, len(data), 4)]
        self.x = self.w[0] | (self.w[1] << 32)
        self.p = 2

    def _renorm(self):
        if self.x < RANS64_L:
            self.x = (self.x << 32) | self.w[self.p]
            self.p += 1

    def _bits(self, n):
        v = self.x & ((1 << n) - 1)
        self.x >>= n
        self._renorm()
        return v

    def decode_stream(self, indexes, cdfs, cdf_lengths, offsets):
        out = []
        mask = (1 << PRECISION) - 1
        for ci in indexes:
            cdf = cdfs[ci]
            max_value = cdf_lengths[ci] - 2
            cum = self.x & mask
            s = 0
            while cdf[s + 1] <= cum:  # first entry > cum, minus one
                s += 1
            start, freq = cdf[s], cdf[s + 1] - cdf[s]
            self.x = freq * (self.x >> PRECISION) + (self.x & mask) - start
            self._renorm()
            value = s
            if value == max_value:
                val = self._bits(BYPASS_PRECISION)
                n_bypass = val
                while val == MAX_BYPASS_VAL:
                    val = self._bits(BYPASS_PRECISION)
                    n_bypass += val
                raw = 0
                for j in range(n_bypass):
                    raw |= self._bits(BYPASS_PRECISION) << (j * BYPASS_PRECISION)
                value = raw >> 1
                if raw & 1:
                    value = -value - 1
                else:
                    value += max_value
            out.append(value + offsets[ci])
        return out


def decode_with_indexes(data, indexes, cdfs, cdf_lengths, offsets):
    return Decoder(data).decode_stream(indexes, cdfs, cdf_lengths, offsets)


def ideal_bits(symbols, indexes, cdfs, cdf_lengths, offsets):
    """Sum of -log2(freq / 2^16) (+4 bits per bypass digit): what the stream length must be close to."""
    bits = 0.0
    for start, rng, bypass in _rans_symbols(symbols, indexes, cdfs, cdf_lengths, offsets):
        bits += BYPASS_PRECISION if bypass else PRECISION - math.log2(rng)
    return bits
