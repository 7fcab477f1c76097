"""Device-side decode of chunked NetCDF-4 / HDF5 variables (csrc/nfx_inflate.cu, C ABI nfx_h5_decode_chunks).

Parity bar: the bytes the GPU produces are identical to zlib.decompress + host unshuffle + numpy placement (what
netCDF4 / HDF5 do for the reference, field.py:22-35,149; subsetNEMO.py:78 writes zlib=True)."""
import os
import zlib

import numpy
import pytest

pytestmark = pytest.mark.gpu


def _payloads(rng, n):
    """byte strings of length n with different statistics: what shuffled ocean fields look like and the corner cases
    of DEFLATE (stored blocks, long matches, short periods, all-literal data)"""
    out = {}
    out['random'] = rng.integers(0, 256, n, dtype=numpy.uint8).tobytes()
    out['zeros'] = bytes(n)
    out['period3'] = (b'abc' * (n // 3 + 1))[:n]
    out['period40'] = (bytes(range(40)) * (n // 40 + 1))[:n]
    x = numpy.cumsum(rng.normal(size=max(n // 4, 1))).astype('<f4')
    sh = numpy.frombuffer(x.tobytes(), numpy.uint8).reshape(-1, 4).T.tobytes()
    out['shuffled_f32'] = (sh * 2)[:n] if len(sh) < n else sh[:n]
    out['few_symbols'] = rng.choice(numpy.frombuffer(b'\x00\x00\x00\x3f\x80\x7f', numpy.uint8), n).tobytes()
    return out


def _compress(raw, level, strategy):
    c = zlib.compressobj(level, zlib.DEFLATED, 15, 9, strategy)
    return c.compress(raw) + c.flush()


def _decode_bytes(gpu, streams, n, filters=1):
    """every stream inflates to n bytes: one call, dst = (nchunks * n) bytes"""
    import torch
    offs, pos = [], 0
    for s in streams:
        offs.append(pos)
        pos += (len(s) + 15) & ~15
    host = numpy.zeros(pos + 4096, numpy.uint8)
    for o, s in zip(offs, streams):
        host[o:o + len(s)] = numpy.frombuffer(s, numpy.uint8)
    comp = torch.from_numpy(host).cuda()
    dst = torch.full((len(streams) * n,), 0xEE, dtype=torch.uint8, device='cuda')
    st = gpu.h5DecodeChunks(comp, offs, [len(s) for s in streams], filters, [n], dst,
                            [[k * n] for k in range(len(streams))])
    return dst.cpu().numpy().reshape(len(streams), n), st


@pytest.mark.parametrize('n', [1, 31, 32, 33, 257, 4096, 65536, 300001])
def test_inflate_is_bit_identical_to_zlib(gpu, n):
    """stored / fixed / dynamic blocks, every zlib level and strategy, streams of several blocks"""
    rng = numpy.random.default_rng(n)
    raws, streams = [], []
    for name, raw in _payloads(rng, n).items():
        for level in (0, 1, 4, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY, zlib.Z_FILTERED):
                if level == 0 and strategy != zlib.Z_DEFAULT_STRATEGY:
                    continue
                s = _compress(raw, level, strategy)
                assert zlib.decompress(s) == raw
                raws.append(raw)
                streams.append(s)
    got, st = _decode_bytes(gpu, streams, n)
    assert (st == 0).all()
    for k, raw in enumerate(raws):
        assert got[k].tobytes() == raw, f'stream {k} of {len(raws)} differs (n = {n})'


def test_inflate_rejects_damaged_streams(gpu):
    """a damaged stream is reported (status + NFX_E_INVALID) and nothing is written outside the chunk"""
    from nemoflux_b200 import _lib
    rng = numpy.random.default_rng(5)
    n = 20000
    raw = _payloads(rng, n)['shuffled_f32']
    good = _compress(raw, 6, zlib.Z_DEFAULT_STRATEGY)
    bad = {}
    b = bytearray(good)
    b[0] = 0x79
    bad['header'] = bytes(b)
    b = bytearray(good)
    b[len(b) // 2] ^= 0x5A
    bad['body'] = bytes(b)
    bad['truncated'] = good[:len(good) // 2]
    b = bytearray(good)
    b[-1] ^= 1
    bad['adler'] = bytes(b)
    bad['too_long'] = _compress(raw + b'xyz', 6, zlib.Z_DEFAULT_STRATEGY)
    bad['too_short'] = _compress(raw[:-5], 6, zlib.Z_DEFAULT_STRATEGY)
    bad['not_zlib'] = bytes(rng.integers(0, 256, 600, dtype=numpy.uint8))
    for name, s in bad.items():
        with pytest.raises(_lib.NemofluxGpuError) as e:
            _decode_bytes(gpu, [good, s, good], n)
        assert 'chunk 1' in str(e.value), (name, str(e.value))
    got, st = _decode_bytes(gpu, [good, good], n)             # the library is usable after the errors
    assert (st == 0).all() and got[0].tobytes() == raw and got[1].tobytes() == raw


@pytest.mark.parametrize('dtype', ['<f4', '<f8', '>f4', '>f8'])
@pytest.mark.parametrize('filters', [(False, False), (True, False), (True, True), (9, True), (1, True)])
def test_device_reader_matches_host_decode(gpu, tmp_path, dtype, filters):
    """chunked variables written by tests/h5build.py (edge chunks hang over the array, two B-tree levels): every block
    read through H5DeviceReader equals h5lite's host decode (zlib + numpy unshuffle) bit for bit, both byte orders"""
    import h5build
    import torch
    from nemoflux_b200 import h5lite
    deflate, shuffle = filters
    rng = numpy.random.default_rng(11)
    data = numpy.cumsum(rng.normal(size=(5, 7, 33, 50)), axis=-1).astype(dtype)
    data[:, 3:, :10, :] = 1.0e20                                          # a block of fill values compresses to long runs
    path = str(tmp_path / 'uo.nc')
    h5build.write(path, {'uo': dict(data=data, chunks=(1, 2, 16, 32), deflate=deflate, shuffle=shuffle, attrs={})})
    f = h5lite.File(path)
    ds = f.datasets['uo']
    rd = gpu.H5DeviceReader(ds, 'cuda')
    for starts, stops in (([0, 0, 0, 0], [5, 7, 33, 50]), ([1, 0, 0, 0], [3, 7, 33, 50]), ([4, 2, 5, 7], [5, 6, 30, 41])):
        host = ds.read_region(starts, stops)
        dev = rd.read(starts, stops).cpu().numpy()
        want = host.astype(host.dtype.newbyteorder('='))
        assert dev.dtype == want.dtype and dev.shape == want.shape
        assert numpy.array_equal(dev.view(numpy.uint8), want.view(numpy.uint8)), (starts, stops)
        assert numpy.array_equal(want, data[tuple(slice(a, b) for a, b in zip(starts, stops))].astype(want.dtype))
    assert rd.bytes_decoded > 0 and (not deflate or rd.bytes_compressed < rd.bytes_decoded)
    f.close()
