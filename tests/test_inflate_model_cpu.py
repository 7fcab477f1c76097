"""CPU checks of the device DEFLATE decoder's control rules through their Python model (tests/inflate_model.py).

The model restates csrc/nfx_inflate.cu::k_inflate's window / budget / flush rules with the kernel's constants and asserts
the invariants the kernel relies on (reader inside the loaded input window, unflushed output never overwritten, far
matches only read flushed bytes).  zlib is the checker for the bytes.  The last two tests show what the model is for:
a budget rule that cannot make progress (the round-2 experiment that hung a GPU call, profiles/r2_fused_experiments.md
section 6) and one that lets the reader outrun the input window both fail here, in milliseconds, without a GPU.
"""
import zlib

import numpy as np
import pytest

from inflate_model import K_FLUSH, K_NEAR, Model, Stuck


def _payloads():
    rng = np.random.default_rng(20261018)
    f = np.cumsum(rng.standard_normal(40000)).astype('<f4')
    out = {
        'empty': b'',
        'one byte': b'a',
        'zeros': bytes(70000),
        'random': rng.integers(0, 256, 50000, dtype=np.uint8).tobytes(),
        'f32': f.tobytes(),
        'f32 shuffled': f.view(np.uint8).reshape(-1, 4).T.copy().tobytes(),
        'text': (b'the quick brown fox jumps over the lazy dog. ' * 2000)[:54321],
        'masked field': np.where(rng.random(30000) < 0.6, 0, rng.standard_normal(30000)).astype('<f4').tobytes(),
    }
    # matches on both sides of the near / far split, and at the longest distance of the format
    block = rng.integers(0, 256, 700, dtype=np.uint8).tobytes()
    far = bytearray()
    for gap in (K_NEAR - 700 - 3, K_NEAR - 700, K_NEAR - 700 + 1, 3000, 8000, 32768 - 700):
        far += block + rng.integers(0, 256, max(gap, 0), dtype=np.uint8).tobytes() + block
    out['near and far matches'] = bytes(far)
    return out


PAYLOADS = _payloads()


def _streams(raw):
    for level in (0, 1, 6, 9):
        yield 'level %d' % level, zlib.compress(raw, level)
    for name, strategy in (('fixed', zlib.Z_FIXED), ('huffman only', zlib.Z_HUFFMAN_ONLY), ('rle', zlib.Z_RLE)):
        c = zlib.compressobj(6, zlib.DEFLATED, 15, 9, strategy)
        yield name, c.compress(raw) + c.flush()
    # many blocks of all three kinds in one stream: full flushes put empty stored blocks between Huffman blocks
    c = zlib.compressobj(6)
    z = b''
    for k in range(0, len(raw), 3001):
        z += c.compress(raw[k:k + 3001]) + c.flush(zlib.Z_FULL_FLUSH if (k // 3001) % 2 else zlib.Z_SYNC_FLUSH)
    yield 'flushed every 3001 bytes', z + c.flush()
    c = zlib.compressobj(1, zlib.DEFLATED, 9)   # 512-byte history
    yield 'small history', c.compress(raw) + c.flush()


@pytest.mark.parametrize('name', sorted(PAYLOADS))
def test_model_inflates_like_zlib(name):
    raw = PAYLOADS[name]
    for kind, z in _streams(raw):
        assert zlib.decompress(z) == raw
        m = Model(z, len(raw))
        assert m.run() == raw, (name, kind)
        # unflushed bytes in the 4 KB window: a flush piece of literals plus the longest match in a Huffman block, the
        # rest of a piece plus a whole stored piece in a stored block -- far from the 4096 that would overwrite them
        assert m.max_unflushed <= max(K_FLUSH + 258, 2 * K_FLUSH - 1), (name, kind, m.max_unflushed)


def test_model_rejects_what_the_kernel_rejects():
    raw = PAYLOADS['f32']
    z = zlib.compress(raw, 6)
    for bad, size in ((z[:len(z) // 2], len(raw)),           # input ends early
                      (z, len(raw) - 1000), (z, len(raw) + 1),   # the chunk is smaller / larger than the stream says
                      (b'\x78\x9d' + z[2:], len(raw)),          # header checksum
                      (z[:2] + b'\x07' + z[3:], len(raw))):     # block type 3
        with pytest.raises(ValueError):
            Model(bad, size).run()
    # every single-byte corruption of a short stream ends in an error or in other bytes, never in a loop or an
    # invariant violation (the Adler-32 pass catches the "other bytes" on the device)
    short = zlib.compress(PAYLOADS['text'][:3000] + PAYLOADS['random'][:500], 9)
    n = 3500
    for k in range(2, len(short) - 4):
        bad = short[:k] + bytes([short[k] ^ 0x55]) + short[k + 1:]
        try:
            Model(bad, n).run()
        except ValueError:
            pass


def test_a_budget_rule_that_cannot_progress_is_caught():
    """Literals two at a time with one shared budget: a budget of 1 rounds down to 0, the round asks for housekeeping,
    housekeeping has nothing to flush (1023 < 1024 bytes) and the next round gets the same budget.  It takes a match of
    odd length to reach the odd budget, so streams of literals alone (and the small test chunks) pass."""
    rule = lambda o, fl, ld, pos: min(K_FLUSH - (o - fl), (ld - 24 - pos) >> 2) & ~1
    raw = PAYLOADS['random']
    assert Model(zlib.compress(raw, 6), len(raw), literal_budget_rule=rule).run() == raw
    for name in ('f32', 'f32 shuffled', 'text'):
        raw = PAYLOADS[name]
        with pytest.raises(Stuck):
            Model(zlib.compress(raw, 6), len(raw), literal_budget_rule=rule).run()


def test_a_budget_rule_that_ignores_the_input_window_is_caught():
    """A flush piece of 1024 literals with 9-bit codes needs 1152 bytes of input; housekeeping only guarantees 1024.
    Half zeros (1-bit code) and half other bytes (9-bit codes), in runs longer than two flush pieces."""
    rng = np.random.default_rng(5)
    raw = b''.join(bytes(2500) + rng.integers(1, 256, 2500, dtype=np.uint8).tobytes() for _ in range(12))
    c = zlib.compressobj(6, zlib.DEFLATED, 15, 9, zlib.Z_HUFFMAN_ONLY)
    z = c.compress(raw) + c.flush()
    assert Model(z, len(raw)).run() == raw
    with pytest.raises(AssertionError, match='loaded window'):
        Model(z, len(raw), literal_budget_rule=lambda o, fl, ld, pos: K_FLUSH - (o - fl)).run()
