"""Test helper: writes a small HDF5 file in the OLDEST on-disk flavour (superblock 0, symbol-table root group with a
local heap and a version-1 B-tree, version-1 object headers, data layout message version 3), by hand from the HDF5
file format specification -- the counterpart of the reference's data/sa/T.nc (superblock 0 with version-2 object
headers and dense links), so that both generations of structures in nemoflux_b200/h5lite.py see a file.
Datasets are float arrays, contiguous or chunked (optionally shuffle + deflate), with numeric / string attributes."""
import struct
import zlib

import numpy

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b):
    return b + b'\0' * (-len(b) % 8)


def _msg(mtype, data):
    data = _pad8(data)
    return struct.pack('<HHBBBB', mtype, len(data), 0, 0, 0, 0) + data


def _dataspace(shape):
    return struct.pack('<BBBB4x', 1, len(shape), 0, 0) + b''.join(struct.pack('<Q', n) for n in shape)


def _datatype(dt):
    dt = numpy.dtype(dt)
    be = 1 if dt.byteorder == '>' else 0
    if dt.kind == 'f':
        if dt.itemsize == 4:
            return struct.pack('<BBBBI', 0x11, 0x20 | be, 31, 0, 4) + struct.pack('<HHBBBBI', 0, 32, 23, 8, 0, 23, 127)
        return struct.pack('<BBBBI', 0x11, 0x20 | be, 63, 0, 8) + struct.pack('<HHBBBBI', 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind in 'iu':
        return struct.pack('<BBBBI', 0x10, be | (0x08 if dt.kind == 'i' else 0), 0, 0, dt.itemsize) + \
            struct.pack('<HH', 0, 8 * dt.itemsize)
    raise ValueError(dt)


def _attribute(name, value):
    if isinstance(value, str):
        raw = value.encode() + b'\0'
        dtype = struct.pack('<BBBBI', 0x13, 0, 0, 0, len(raw))
        space = _dataspace(())
    else:
        a = numpy.atleast_1d(numpy.asarray(value))
        raw = a.tobytes()
        dtype = _datatype(a.dtype)
        space = _dataspace(a.shape if numpy.ndim(value) else ())
    nm = name.encode() + b'\0'
    return struct.pack('<BBHHH', 1, 0, len(nm), len(dtype), len(space)) + _pad8(nm) + _pad8(dtype) + _pad8(space) + raw


def write(path, variables):
    """variables: {name: dict(data=array, chunks=None or tuple, deflate=False, shuffle=False, attrs={})}"""
    buf = bytearray(b'\0' * 96)                                  # superblock, filled in at the end

    def put(b):
        addr = len(buf)
        buf.extend(_pad8(bytes(b)))
        return addr

    def chunk_tree(data, cdims, deflate, shuffle):
        rank = data.ndim
        items = []
        grid = [range(0, n, c) for n, c in zip(data.shape, cdims)]
        for offs in numpy.stack(numpy.meshgrid(*grid, indexing='ij'), -1).reshape(-1, rank):
            chunk = numpy.zeros(cdims, data.dtype)
            sub = data[tuple(slice(o, o + c) for o, c in zip(offs, cdims))]
            chunk[tuple(slice(0, m) for m in sub.shape)] = sub
            raw = chunk.tobytes()
            if shuffle:
                raw = numpy.frombuffer(raw, numpy.uint8).reshape(-1, data.dtype.itemsize).T.tobytes()
            if deflate:
                raw = zlib.compress(raw, deflate if type(deflate) is int else 5)
            items.append((tuple(int(o) for o in offs), len(raw), put(raw)))

        def node(level, entries, last_key):
            b = b'TREE' + bytes([1, level]) + struct.pack('<H', len(entries)) + struct.pack('<QQ', UNDEF, UNDEF)
            for offs, nbytes, child in entries:
                b += struct.pack('<II', nbytes, 0) + b''.join(struct.pack('<Q', o) for o in offs) + struct.pack('<Q', 0)
                b += struct.pack('<Q', child)
            b += struct.pack('<II', 0, 0) + b''.join(struct.pack('<Q', o) for o in last_key) + struct.pack('<Q', 0)
            return put(b)

        end = tuple(int(n) for n in data.shape)
        if len(items) <= 4:
            return node(0, items, end)
        leaves = []
        for i in range(0, len(items), 4):                        # two levels: leaves of up to 4 chunks
            part = items[i:i + 4]
            nxt = items[i + 4][0] if i + 4 < len(items) else end
            leaves.append((part[0][0], 0, node(0, part, nxt)))
        return node(1, leaves, end)

    heap_data = bytearray(b'\0' * 8)
    symbols = []
    for name, spec in sorted(variables.items()):
        data = numpy.ascontiguousarray(spec['data'])
        msgs = _msg(0x01, _dataspace(data.shape)) + _msg(0x03, _datatype(data.dtype))
        msgs += _msg(0x05, struct.pack('<BBBB', 2, 2, 0, 0))     # fill value message v2, undefined
        chunks = spec.get('chunks')
        if chunks:
            filt = []
            if spec.get('shuffle'):
                filt.append((2, [data.dtype.itemsize]))
            if spec.get('deflate'):
                filt.append((1, [5]))
            if filt:
                body = struct.pack('<BB6x', 1, len(filt))
                for fid, cd in filt:
                    body += struct.pack('<HHHH', fid, 0, 1, len(cd)) + b''.join(struct.pack('<I', c) for c in cd)
                    if len(cd) % 2:
                        body += b'\0' * 4
                msgs += _msg(0x0B, body)
            tree = chunk_tree(data, tuple(chunks), spec.get('deflate', False), spec.get('shuffle', False))
            lay = struct.pack('<BBB', 3, 2, data.ndim + 1) + struct.pack('<Q', tree)
            lay += b''.join(struct.pack('<I', c) for c in chunks) + struct.pack('<I', data.dtype.itemsize)
        else:
            addr = put(data.tobytes())
            lay = struct.pack('<BB', 3, 1) + struct.pack('<QQ', addr, data.nbytes)
        msgs += _msg(0x08, lay)
        for k, v in spec.get('attrs', {}).items():
            msgs += _msg(0x0C, _attribute(k, v))
        nmsg = 4 + (1 if chunks and (spec.get('shuffle') or spec.get('deflate')) else 0) + len(spec.get('attrs', {}))
        oh = put(struct.pack('<BBHII4x', 1, 0, nmsg, 1, len(msgs)) + msgs)
        noff = len(heap_data)
        heap_data.extend(_pad8(name.encode() + b'\0'))
        symbols.append((noff, oh))
    heap_seg = put(heap_data)
    heap = put(b'HEAP' + struct.pack('<B3x', 0) + struct.pack('<QQQ', len(heap_data), UNDEF, heap_seg))
    snod = b'SNOD' + struct.pack('<BBH', 1, 0, len(symbols))
    for noff, oh in symbols:
        snod += struct.pack('<QQII16x', noff, oh, 0, 0)
    snod_addr = put(snod)
    btree = put(b'TREE' + bytes([0, 0]) + struct.pack('<H', 1) + struct.pack('<QQ', UNDEF, UNDEF) +
                struct.pack('<QQQ', 0, snod_addr, symbols[-1][0]))
    root_msgs = _msg(0x11, struct.pack('<QQ', btree, heap))
    root = put(struct.pack('<BBHII4x', 1, 0, 1, 1, len(root_msgs)) + root_msgs)
    sb = b'\x89HDF\r\n\x1a\n' + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack('<HHI', 4, 16, 0)
    sb += struct.pack('<QQQQ', 0, UNDEF, len(buf), UNDEF)
    sb += struct.pack('<QQII', 0, root, 1, 0) + struct.pack('<QQ', btree, heap)
    buf[:len(sb)] = sb
    with open(path, 'wb') as f:
        f.write(bytes(buf))


# ---------------------------------------------------------------------------------------------------------------
# the NEWER flavour: superblock 2, version-2 object headers (times, creation order, a continuation block), link
# messages in the root header, dataspace v2, fill value v3, filter pipeline v2, attribute messages v3 (one
# variable-length string through a global heap), and -- for variables with `dense_attrs` -- attributes in dense storage:
# a fractal heap whose root is an INDIRECT block (small starting blocks) indexed by a version-2 B-tree
# ---------------------------------------------------------------------------------------------------------------
def _msg2(mtype, data, corder=0):
    return struct.pack('<BHB', mtype, len(data), 0) + struct.pack('<H', corder) + data


def _dataspace2(shape):
    return struct.pack('<BBBB', 2, len(shape), 0, 1 if len(shape) else 0) + b''.join(struct.pack('<Q', n) for n in shape)


def _attribute3(name, value, put_gcol):
    if isinstance(value, str) and value.startswith('vlen:'):
        text = value[5:].encode()
        gaddr, idx = put_gcol(text)
        raw = struct.pack('<IQI', len(text), gaddr, idx)
        dtype = struct.pack('<BBBBI', 0x19, 0x01, 0x01, 0, 16) + struct.pack('<BBBBI', 0x13, 0, 0, 0, 1)
        space = _dataspace2(())
    elif isinstance(value, str):
        raw = value.encode() + b'\0'
        dtype = struct.pack('<BBBBI', 0x13, 0, 0, 0, len(raw))
        space = _dataspace2(())
    else:
        a = numpy.atleast_1d(numpy.asarray(value))
        raw = a.tobytes()
        dtype = _datatype(a.dtype)
        space = _dataspace2(a.shape if numpy.ndim(value) else ())
    nm = name.encode() + b'\0'
    return struct.pack('<BBHHHB', 3, 0, len(nm), len(dtype), len(space), 0) + nm + dtype + space + raw


def write_new(path, variables):
    """same `variables` as write(); extra keys: dense_attrs=True puts the attributes into a fractal heap"""
    buf = bytearray(b'\0' * 48)

    def put(b, align=8):
        buf.extend(b'\0' * (-len(buf) % align))
        addr = len(buf)
        buf.extend(bytes(b))
        return addr

    gcol_objs = []

    def put_gcol(data):                                           # all strings go into ONE collection, written last
        gcol_objs.append(data)
        return GCOL_ADDR[0], len(gcol_objs)

    GCOL_ADDR = [0]
    # the collection's address must be known before the attributes are encoded: reserve room for it first
    GCOL_SIZE = 4096
    GCOL_ADDR[0] = put(b'\0' * GCOL_SIZE)

    def ohdr(messages, with_continuation=False):
        first, rest = (messages[:2], messages[2:]) if with_continuation and len(messages) > 3 else (messages, [])
        body = b''.join(first)
        if rest:
            cont = b'OCHK' + b''.join(rest) + b'\0\0\0\0'
            caddr = put(cont)
            body += _msg2(0x10, struct.pack('<QQ', caddr, len(cont)))
        body += _msg2(0x00, b'\0' * 6)                            # a NIL message and a short gap before the checksum
        prefix = b'OHDR' + bytes([2, 0x02 | 0x04 | 0x20]) + struct.pack('<IIII', 1, 2, 3, 4)
        return put(prefix + struct.pack('<I', len(body) + 3) + body + b'\0' * 3 + b'\0\0\0\0')

    def dense_attributes(attrs):
        width, start, max_direct, obytes = 4, 256, 4096, 2
        hdr_len = 5 + 8 + obytes
        blocks, records = [], []                                   # blocks: [bytearray]; heap offset = sum of sizes before
        sizes = [start] * (2 * width) + [2 * start] * width
        cur, pos, base = 0, hdr_len, 0
        blocks.append(bytearray(sizes[0]))
        for n, (k, v) in enumerate(attrs.items()):
            enc = _attribute3(k, v, put_gcol)
            if pos + len(enc) > sizes[cur]:
                base += sizes[cur]
                cur += 1
                blocks.append(bytearray(sizes[cur]))
                pos = hdr_len
            blocks[cur][pos:pos + len(enc)] = enc
            hid = bytes([0]) + struct.pack('<H', base + pos) + struct.pack('<H', len(enc)) + b'\0' * 3
            records.append(hid + bytes([0]) + struct.pack('<II', n, 0x1000 + n))
            pos += len(enc)
        heap_addr = put(b'\0' * 146)
        baddrs, off = [], 0
        for i, blk in enumerate(blocks):
            blk[:hdr_len] = b'FHDB' + bytes([0]) + struct.pack('<Q', heap_addr) + struct.pack('<H', off)
            baddrs.append(put(blk))
            off += sizes[i]
        rows = (len(blocks) + width - 1) // width
        table = b''.join(struct.pack('<Q', baddrs[i] if i < len(baddrs) else UNDEF) for i in range(rows * width))
        root = put(b'FHIB' + bytes([0]) + struct.pack('<Q', heap_addr) + struct.pack('<H', 0) + table + b'\0\0\0\0')
        hdr = b'FRHP' + bytes([0]) + struct.pack('<HHB', 8, 0, 0) + struct.pack('<I', 4096)
        hdr += struct.pack('<Q', 0) + struct.pack('<Q', UNDEF) + struct.pack('<Q', 0) + struct.pack('<Q', UNDEF)
        hdr += struct.pack('<QQQQ', off, off, 0, len(records)) + struct.pack('<QQQQ', 0, 0, 0, 0)
        hdr += struct.pack('<HQQHH', width, start, max_direct, 16, 1) + struct.pack('<QH', root, rows) + b'\0\0\0\0'
        assert len(hdr) == 146, len(hdr)
        buf[heap_addr:heap_addr + 146] = hdr
        leaf = put(b'BTLF' + bytes([0, 8]) + b''.join(records) + b'\0\0\0\0')
        bt = put(b'BTHD' + bytes([0, 8]) + struct.pack('<IHHBB', 512, 17, 0, 100, 40) +
                 struct.pack('<QHQ', leaf, len(records), len(records)) + b'\0\0\0\0')
        return struct.pack('<BB', 0, 0) + struct.pack('<QQ', heap_addr, bt)

    links = []
    for n, (name, spec) in enumerate(sorted(variables.items())):
        data = numpy.ascontiguousarray(spec['data'])
        msgs = [_msg2(0x01, _dataspace2(data.shape)), _msg2(0x03, _datatype(data.dtype)),
                _msg2(0x05, struct.pack('<BB', 3, 0x09))]        # fill value v3: allocate early, undefined
        chunks = spec.get('chunks')
        if chunks:
            filt = ([(2, [data.dtype.itemsize])] if spec.get('shuffle') else []) + ([(1, [5])] if spec.get('deflate') else [])
            if filt:
                body = struct.pack('<BB', 2, len(filt))
                for fid, cd in filt:
                    body += struct.pack('<HHH', fid, 1, len(cd)) + b''.join(struct.pack('<I', c) for c in cd)
                msgs.append(_msg2(0x0B, body))
            # the chunk B-tree writer of write() works on its own buffer: reuse it through a scratch file image
            tree = _chunk_tree_into(buf, data, tuple(chunks), spec.get('deflate', False), spec.get('shuffle', False))
            lay = struct.pack('<BBB', 3, 2, data.ndim + 1) + struct.pack('<Q', tree)
            lay += b''.join(struct.pack('<I', c) for c in chunks) + struct.pack('<I', data.dtype.itemsize)
        else:
            lay = struct.pack('<BB', 3, 1) + struct.pack('<QQ', put(data.tobytes()), data.nbytes)
        msgs.append(_msg2(0x08, lay))
        attrs = spec.get('attrs', {})
        if spec.get('dense_attrs'):
            msgs.append(_msg2(0x15, dense_attributes(attrs)))
        else:
            for i, (k, v) in enumerate(attrs.items()):
                msgs.append(_msg2(0x0C, _attribute3(k, v, put_gcol), corder=i))
        addr = ohdr(msgs, with_continuation=True)
        nm = name.encode()
        # link message: version 1, flags = creation order (0x04) + charset (0x10), 1-byte name length
        links.append(_msg2(0x06, struct.pack('<BB', 1, 0x14) + struct.pack('<Q', n) + bytes([1, len(nm)]) + nm +
                           struct.pack('<Q', addr), corder=n))
    root = ohdr(links + [_msg2(0x0C, _attribute3('Conventions', 'vlen:CF-1.6', put_gcol))], with_continuation=True)
    col = b'GCOL' + bytes([1, 0, 0, 0]) + struct.pack('<Q', GCOL_SIZE)
    for i, d in enumerate(gcol_objs):
        col += struct.pack('<HHIQ', i + 1, 1, 0, len(d)) + _pad8(d)
    assert len(col) + 16 <= GCOL_SIZE
    buf[GCOL_ADDR[0]:GCOL_ADDR[0] + len(col)] = col
    sb = b'\x89HDF\r\n\x1a\n' + bytes([2, 8, 8, 0]) + struct.pack('<QQQQ', 0, UNDEF, len(buf), root) + b'\0\0\0\0'
    buf[:len(sb)] = sb
    with open(path, 'wb') as f:
        f.write(bytes(buf))


def _chunk_tree_into(buf, data, cdims, deflate, shuffle):
    """version-1 chunk B-tree (one leaf level, then one internal level when there are more than 4 chunks)"""
    def put(b):
        buf.extend(b'\0' * (-len(buf) % 8))
        addr = len(buf)
        buf.extend(bytes(b))
        return addr
    rank = data.ndim
    items = []
    grid = [range(0, n, c) for n, c in zip(data.shape, cdims)]
    for offs in numpy.stack(numpy.meshgrid(*grid, indexing='ij'), -1).reshape(-1, rank):
        chunk = numpy.zeros(cdims, data.dtype)
        sub = data[tuple(slice(o, o + c) for o, c in zip(offs, cdims))]
        chunk[tuple(slice(0, m) for m in sub.shape)] = sub
        raw = chunk.tobytes()
        if shuffle:
            raw = numpy.frombuffer(raw, numpy.uint8).reshape(-1, data.dtype.itemsize).T.tobytes()
        if deflate:
            raw = zlib.compress(raw, deflate if type(deflate) is int else 5)
        items.append((tuple(int(o) for o in offs), len(raw), put(raw)))

    def node(level, entries, last_key):
        b = b'TREE' + bytes([1, level]) + struct.pack('<H', len(entries)) + struct.pack('<QQ', UNDEF, UNDEF)
        for offs, nbytes, child in entries:
            b += struct.pack('<II', nbytes, 0) + b''.join(struct.pack('<Q', o) for o in offs) + struct.pack('<Q', 0)
            b += struct.pack('<Q', child)
        b += struct.pack('<II', 0, 0) + b''.join(struct.pack('<Q', o) for o in last_key) + struct.pack('<Q', 0)
        return put(b)
    end = tuple(int(n) for n in data.shape)
    if len(items) <= 4:
        return node(0, items, end)
    leaves = []
    for i in range(0, len(items), 4):
        part = items[i:i + 4]
        leaves.append((part[0][0], 0, node(0, part, items[i + 4][0] if i + 4 < len(items) else end)))
    return node(1, leaves, end)
