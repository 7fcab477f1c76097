"""A line-by-line Python model of the CONTROL LOGIC of csrc/nfx_inflate.cu::k_inflate -- test infrastructure.

Not an oracle of the reference (zlib is that) and not a product path: the model restates, with the kernel's own
constants, the parts of the device decoder that are easy to get wrong and expensive to debug on a GPU --

  * the bit reader that moves aligned 32-bit words from a 2 KB ring of the input into a 64-bit buffer,
  * the ring's refill rule (a half is overwritten only when the reader is past it),
  * the tight literal loop with ONE budget for the flush piece and the input window,
  * the 4 KB output window, its 1 KB flush pieces, and the near / far split of the match copies,
  * the deferred overrun check and the stored-block restart --

so that a change of those rules can be run against zlib on the CPU first (tests/test_host_cpu.py).  Every loop carries
an iteration cap: a rule that stops making progress fails the test instead of hanging a GPU call.
Huffman table construction is plain canonical decoding here (the kernel's lookup tables are exercised on the GPU).
"""
K_IN_RING = 2048
K_IN_HALF = K_IN_RING // 2
K_WIN = 4096
K_FLUSH = 1024
K_NEAR = K_WIN - 258 - K_FLUSH - 258
K_LIT_BITS = 10   # codes up to this length come out of the kernel's lookup table; longer ones leave the tight loop

LEN_TAB = [(3, 0), (4, 0), (5, 0), (6, 0), (7, 0), (8, 0), (9, 0), (10, 0), (11, 1), (13, 1), (15, 1), (17, 1), (19, 2), (23, 2),
           (27, 2), (31, 2), (35, 3), (43, 3), (51, 3), (59, 3), (67, 4), (83, 4), (99, 4), (115, 4), (131, 5), (163, 5),
           (195, 5), (227, 5), (258, 0)]
DIST_TAB = [(1, 0), (2, 0), (3, 0), (4, 0), (5, 1), (7, 1), (9, 2), (13, 2), (17, 3), (25, 3), (33, 4), (49, 4), (65, 5), (97, 5),
            (129, 6), (193, 6), (257, 7), (385, 7), (513, 8), (769, 8), (1025, 9), (1537, 9), (2049, 10), (3073, 10),
            (4097, 11), (6145, 11), (8193, 12), (12289, 12), (16385, 13), (24577, 13)]
CLEN_ORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]


class Stuck(Exception):
    """a loop of the model exceeded its iteration cap: the rule under test makes no progress"""


class Model(object):

    def __init__(self, stream, out_size, literal_budget_rule=None):
        self.inp = bytes(stream)
        self.in_size = len(stream)
        self.out_size = out_size
        self.ring = bytearray(K_IN_RING)
        self.loaded = 0
        self.win = bytearray(K_WIN)
        self.out = bytearray(out_size)
        self.opos = self.flushed = 0
        self.buf = self.cnt = self.pos = 0
        # budget of the tight loop given (opos, flushed, loaded, pos): the rule a kernel change would touch
        self.budget_rule = literal_budget_rule or (lambda o, fl, ld, pos: min(K_FLUSH - (o - fl), (ld - 24 - pos) >> 2))
        self.max_unflushed = 0

    # ---- input ring -------------------------------------------------------------------------------------------
    def load_half(self, base):
        for k in range(K_IN_HALF):
            off = base + k
            self.ring[off & (K_IN_RING - 1)] = self.inp[off] if off < self.in_size else 0

    def top_up(self):
        n = 0
        while self.loaded - self.pos < K_IN_HALF:
            # the half [loaded - K_IN_RING, loaded - K_IN_HALF) is overwritten: the reader must be past it
            assert self.pos >= self.loaded - K_IN_HALF, 'a half of the input ring would be overwritten before it was read'
            self.load_half(self.loaded)
            self.loaded += K_IN_HALF
            n += 1
            if n > 8:
                raise Stuck('top_up')

    def refill(self):
        while self.cnt <= 32:
            assert self.pos % 4 == 0 and self.pos + 4 <= self.loaded, 'the bit reader ran past the loaded window'
            assert self.pos >= self.loaded - K_IN_RING, 'the bit reader reads bytes that were already overwritten'
            w = int.from_bytes(self.ring[self.pos & (K_IN_RING - 1):(self.pos & (K_IN_RING - 1)) + 4], 'little')
            self.buf |= w << self.cnt
            self.cnt += 32
            self.pos += 4

    def take(self, n):
        assert self.cnt >= n
        v = self.buf & ((1 << n) - 1)
        self.buf >>= n
        self.cnt -= n
        return v

    # ---- canonical Huffman ------------------------------------------------------------------------------------
    @staticmethod
    def tables(lens):
        cnt = [0] * 16
        for l in lens:
            cnt[l] += 1
        offs = [0] * 16
        for l in range(1, 15):
            offs[l + 1] = offs[l] + cnt[l]
        sym = [0] * len(lens)
        for s, l in enumerate(lens):
            if l:
                sym[offs[l]] = s
                offs[l] += 1
        left = 1
        for l in range(1, 16):
            left = (left << 1) - cnt[l]
            if left < 0:
                raise ValueError('over-subscribed code')
        return cnt, sym

    def peek_decode(self, tab):
        """(symbol, code length) of the next code without consuming it"""
        cnt, sym = tab
        code = first = index = 0
        bits = self.buf
        for l in range(1, 16):
            assert self.cnt >= l
            code |= bits & 1
            bits >>= 1
            c = cnt[l]
            if code - c < first:
                return sym[index + (code - first)], l
            index += c
            first = (first + c) << 1
            code <<= 1
        raise ValueError('invalid code')

    def decode(self, tab):
        s, l = self.peek_decode(tab)
        self.take(l)
        return s

    # ---- output window ----------------------------------------------------------------------------------------
    def flush_pieces(self):
        n = 0
        while self.opos - self.flushed >= K_FLUSH:
            for p in range(self.flushed, self.flushed + K_FLUSH):
                self.out[p] = self.win[p & (K_WIN - 1)]
            self.flushed += K_FLUSH
            n += 1
            if n > 8:
                raise Stuck('flush')

    def housekeeping(self):
        if self.opos > self.out_size:
            raise ValueError('overrun')
        self.top_up()
        self.flush_pieces()

    def put(self, b):
        self.win[self.opos & (K_WIN - 1)] = b
        self.opos += 1
        self.max_unflushed = max(self.max_unflushed, self.opos - self.flushed)
        assert self.opos - self.flushed <= K_WIN, 'unflushed output was overwritten in the window'

    def copy_match(self, length, dist):
        if dist <= K_NEAR:
            assert dist + length <= K_WIN, 'a near match reads bytes its own output has overwritten in the window'
            for i in range(length):
                self.put(self.win[(self.opos - dist) & (K_WIN - 1)])
        else:   # far: the source must have been flushed to the chunk already
            assert self.opos - dist + length <= self.flushed, 'a far match reads bytes that are not in the chunk yet'
            for i in range(length):
                self.put(self.out[self.opos - dist])

    # ---- the decoder ------------------------------------------------------------------------------------------
    def run(self):
        self.load_half(0)
        self.load_half(K_IN_HALF)
        self.loaded = K_IN_RING
        self.refill()
        cmf, flg = self.take(8), self.take(8)
        if (cmf & 15) != 8 or (flg & 32) or ((cmf << 8) | flg) % 31 or self.in_size < 6:
            raise ValueError('bad header')
        last = 0
        blocks = 0
        while not last:
            blocks += 1
            if blocks > 100000:
                raise Stuck('blocks')
            self.top_up()
            self.refill()
            last, btype = self.take(1), self.take(2)
            if btype == 3:
                raise ValueError('bad block')
            if btype == 0:
                self.take(self.cnt & 7)
                self.refill()
                ln, nl = self.take(16), self.take(16)
                if (ln ^ 0xffff) != nl:
                    raise ValueError('bad stored block')
                src = self.pos - self.cnt // 8
                if src + ln > self.in_size:
                    raise ValueError('truncated')
                if self.opos + ln > self.out_size:
                    raise ValueError('overrun')
                done = 0
                while done < ln:
                    piece = min(ln - done, K_FLUSH)
                    for i in range(piece):
                        self.put(self.inp[src + done + i])
                    done += piece
                    self.flush_pieces()
                np_ = src + ln
                base = np_ & ~(K_IN_HALF - 1)
                self.load_half(base)
                self.load_half(base + K_IN_HALF)
                self.loaded = base + K_IN_RING
                self.pos = (np_ & ~3) + 4
                self.cnt = 32 - 8 * (np_ & 3)
                w = int.from_bytes(self.ring[(np_ & ~3) & (K_IN_RING - 1):((np_ & ~3) & (K_IN_RING - 1)) + 4], 'little')
                self.buf = w >> (8 * (np_ & 3))
                continue
            if btype == 1:
                lens = [8] * 144 + [9] * 112 + [7] * 24 + [8] * 8
                dlens = [5] * 30
            else:
                self.refill()
                nlit, ndist, nclen = self.take(5) + 257, self.take(5) + 1, self.take(4) + 4
                if nlit > 286 or ndist > 30:
                    raise ValueError('bad code')
                cl = [0] * 19
                for i in range(nclen):
                    self.refill()
                    cl[CLEN_ORDER[i]] = self.take(3)
                ctab = self.tables(cl)
                ls, prev = [], 0
                start = self.pos
                while len(ls) < nlit + ndist:
                    self.refill()
                    s = self.decode(ctab)
                    if s < 16:
                        prev = s
                        ls.append(s)
                    else:
                        if s == 16:
                            if not ls:
                                raise ValueError('bad code')
                            val, rep = prev, 3 + self.take(2)
                        elif s == 17:
                            val, rep = 0, 3 + self.take(3)
                        else:
                            val, rep = 0, 11 + self.take(7)
                        if len(ls) + rep > nlit + ndist:
                            raise ValueError('bad code')
                        ls += [val] * rep
                        prev = val
                assert self.pos - start <= K_IN_HALF - 64, 'a dynamic header outgrew what the block-start top-up guarantees'
                if ls[256] == 0:
                    raise ValueError('no end of block')
                lens, dlens = ls[:nlit] + [0] * (288 - nlit), ls[nlit:] + [0] * (30 - ndist)
            lit, dis = self.tables(lens), self.tables(dlens)
            need_hk = True
            rounds = 0
            while True:
                rounds += 1
                if rounds > 4 * (self.out_size + self.in_size) + 1000:
                    raise Stuck('symbol loop: no progress (budget rule)')
                if need_hk:
                    self.housekeeping()
                    need_hk = False
                budget = self.budget_rule(self.opos, self.flushed, self.loaded, self.pos)
                sym = -1
                while budget > 0:
                    budget -= 1
                    if self.cnt <= 32:   # ONE word per literal, as in the kernel: the budget counts on it
                        self.refill()
                    s, l = self.peek_decode(lit)
                    if l > K_LIT_BITS:
                        sym = -2
                        break
                    self.take(l)
                    if s < 256:
                        self.put(s)
                        continue
                    sym = s
                    break
                if sym == -2:   # a code longer than the lookup: decoded outside the tight loop
                    self.refill()
                    sym = self.decode(lit)
                    if sym < 256:
                        self.put(sym)
                        sym = -1
                if sym == -1:
                    if self.pos - self.cnt // 8 > self.in_size:
                        raise ValueError('truncated')
                    need_hk = True
                    continue
                if sym == 256:
                    if self.pos - self.cnt // 8 > self.in_size:
                        raise ValueError('truncated')
                    break
                if sym - 257 >= 29:
                    raise ValueError('bad length code')
                self.refill()
                base, extra = LEN_TAB[sym - 257]
                length = base + self.take(extra)
                self.refill()
                d = self.decode(dis)
                if d >= 30:
                    raise ValueError('bad distance code')
                base, extra = DIST_TAB[d]
                dist = base + self.take(extra)
                if dist > self.opos:
                    raise ValueError('bad distance')
                if self.opos + length > self.out_size:
                    raise ValueError('overrun')
                if self.pos - self.cnt // 8 > self.in_size:
                    raise ValueError('truncated')
                self.copy_match(length, dist)
                if self.opos - self.flushed >= K_FLUSH:
                    need_hk = True
        if self.opos != self.out_size:
            raise ValueError('overrun' if self.opos > self.out_size else 'short')
        for p in range(self.flushed, self.opos):
            self.out[p] = self.win[p & (K_WIN - 1)]
        return bytes(self.out)
