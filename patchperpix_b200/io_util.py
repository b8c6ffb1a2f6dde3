"""Containers on either side of the assembly stage.

The reference reads predictions from zarr (written by predict_no_gp.py:243-257 /
decode.py:102-109) or hdf, caches per-block results in a zarr next to the output
(stitch_patch_graph.py:649-669, 338-357) and writes `<sample>.hdf`
(vote_instances.py:542-554, stitch_patch_graph.py:849-894).  zarr and h5py are
optional packages; when zarr is missing, `ZarrLite` below reads and writes the
zarr v2 directory layout itself (C order, chunked, compressor null / zlib / gzip /
zstd / blosc) -- enough for the bundled flylight sample (gzip chunks), for stores
written by this package and for the stores of a predict run.
Blosc frames (what predict_no_gp.py:243-257 writes: zstd, bit-shuffle) are unpacked by
`blosc_decode` below, written from the c-blosc 1.x chunk format; no Blosc encoder
exists in this image, so it is checked against frames built by the tests' own
encoder (tests/test_boundary.py), not against numcodecs output -- with zarr +
numcodecs installed `open_container` uses those and never reaches it.
`.npz` is the always-available result format.
"""
import itertools
import json
import logging
import os
import zlib

import numpy as np

logger = logging.getLogger(__name__)


# ---------------------------------------------------------------------------
# zarr v2 directory stores without the zarr package
# ---------------------------------------------------------------------------
def _decompress(buf, comp, path):
    if comp is None:
        return buf
    cid = comp.get('id')
    if cid == 'zlib':
        return zlib.decompress(buf)
    if cid == 'gzip':
        return zlib.decompress(buf, 16 + zlib.MAX_WBITS)
    if cid == 'zstd':
        return _zstd_decompress(buf)
    if cid == 'blosc':
        return blosc_decode(buf, path)
    raise RuntimeError("%s: chunks are compressed with %r; install zarr + numcodecs to read "
                       "this store (ZarrLite handles null, zlib, gzip, zstd, blosc)" % (path, cid))


_BLOSC_MAX_SPLITS, _BLOSC_MIN_BUFFERSIZE = 16, 128


def _blosc_codec(code, path):
    """compressor format in bits 5-7 of the flags byte -> f(bytes, size) -> bytes."""
    import pyarrow as pa
    if code == 4:
        return lambda b, n: pa.Codec('zstd').decompress(b, decompressed_size=n, asbytes=True)
    if code == 3:
        return lambda b, n: zlib.decompress(b)
    if code == 1:
        return lambda b, n: pa.Codec('lz4_raw').decompress(b, decompressed_size=n, asbytes=True)
    if code == 2:
        return lambda b, n: pa.Codec('snappy').decompress(b, decompressed_size=n, asbytes=True)
    raise RuntimeError("%s: blosc frame with internal codec %d (blosclz) is not supported; "
                       "install zarr + numcodecs" % (path, code))


def _bit_unshuffle(block, typesize):
    """inverse of bitshuffle's bit transpose over whole elements: the input holds, for
    byte j of the element and bit k of that byte, one row of size/8 bytes whose bit i%8
    of byte i/8 (least significant first) belongs to element i.  c-blosc 1.x shuffles
    only blocks whose element count is a multiple of 8 and copies any other block."""
    n = len(block)
    size = n // typesize
    if size % 8 != 0 or size == 0:
        return block
    body = np.frombuffer(block, np.uint8, size * typesize).reshape(typesize, 8, size // 8)
    # one 64-bit word per (element byte j, group of 8 elements): byte k of the word is the
    # group's byte of bit row k; transposing the word's 8x8 bit matrix turns it into the
    # eight element bytes (three masked swaps)
    x = np.ascontiguousarray(body.transpose(0, 2, 1)).view('<u8')[..., 0]
    for sh, msk in ((7, 0x00AA00AA00AA00AA), (14, 0x0000CCCC0000CCCC), (28, 0x00000000F0F0F0F0)):
        t = (x ^ (x >> np.uint64(sh))) & np.uint64(msk)
        x = x ^ t ^ (t << np.uint64(sh))
    out = np.ascontiguousarray(x.view(np.uint8).reshape(typesize, size).T)   # [i][j]
    return out.tobytes() + bytes(block[size * typesize:])


def _byte_unshuffle(block, typesize):
    n = len(block)
    q = n // typesize
    body = np.frombuffer(block, np.uint8, q * typesize).reshape(typesize, q)
    return np.ascontiguousarray(body.T).tobytes() + bytes(block[q * typesize:])


def blosc_decode(buf, path='<buffer>'):
    """one c-blosc 1.x chunk -> bytes.  16-byte header: version, versionlz, flags,
    typesize, nbytes, blocksize, cbytes (u32 little endian); flags: 0x1 byte shuffle,
    0x2 stored uncompressed, 0x4 bit shuffle, 0x10 blocks not split, bits 5-7 the codec.
    Then one i32 start offset per block; a block is `typesize` streams (one per byte of the
    element, when split) or one stream, each an i32 length + data, stored raw when the
    length equals the stream's uncompressed size."""
    buf = bytes(buf)
    if len(buf) < 16:
        raise RuntimeError("%s: truncated blosc chunk" % path)
    flags, typesize = buf[2], buf[3]
    nbytes, blocksize, cbytes = (int.from_bytes(buf[o:o + 4], 'little') for o in (4, 8, 12))
    if cbytes != len(buf):
        raise RuntimeError("%s: blosc header says %d bytes, the chunk has %d" % (path, cbytes, len(buf)))
    if nbytes == 0:
        return b''
    if flags & 0x2:
        return buf[16:16 + nbytes]
    codec = _blosc_codec(flags >> 5, path)
    shuffle = bool(flags & 0x1) and typesize > 1
    bitshuffle = bool(flags & 0x4)
    nblocks = (nbytes + blocksize - 1) // blocksize
    out = []
    for b in range(nblocks):
        bsize = min(blocksize, nbytes - b * blocksize)
        leftover = bsize != blocksize
        split = (not (flags & 0x10) and typesize <= _BLOSC_MAX_SPLITS and
                 blocksize // typesize >= _BLOSC_MIN_BUFFERSIZE and not leftover)
        nsplits = typesize if split else 1
        ne = bsize // nsplits
        pos = int.from_bytes(buf[16 + 4 * b:20 + 4 * b], 'little', signed=True)
        parts = []
        for _ in range(nsplits):
            cb = int.from_bytes(buf[pos:pos + 4], 'little', signed=True)
            pos += 4
            if cb < 0 or pos + cb > len(buf):
                raise RuntimeError("%s: corrupt blosc block %d" % (path, b))
            part = buf[pos:pos + cb] if cb == ne else codec(buf[pos:pos + cb], ne)
            if len(part) != ne:
                raise RuntimeError("%s: blosc block %d unpacked to %d bytes, expected %d"
                                   % (path, b, len(part), ne))
            parts.append(part)
            pos += cb
        block = b''.join(parts)
        if shuffle:
            block = _byte_unshuffle(block, typesize)
        elif bitshuffle and bsize >= typesize:
            block = _bit_unshuffle(block, typesize)
        out.append(block)
    return b''.join(out)


def _zstd_decompress(buf):
    """zstd frame -> bytes through pyarrow's codec (needs the frame's content size)."""
    import pyarrow as pa
    # frame header: magic (4) + descriptor; content size is present when written by
    # numcodecs / zstd's simple API
    fhd = buf[4]
    fcs_flag, single = fhd >> 6, (fhd >> 5) & 1
    pos = 5 + (0 if single else 1) + [0, 1, 2, 4][fhd & 3]
    size_len = [1 if single else 0, 2, 4, 8][fcs_flag]
    if size_len == 0:
        raise RuntimeError("zstd frame without content size")
    size = int.from_bytes(buf[pos:pos + size_len], 'little') + (256 if size_len == 2 else 0)
    return pa.Codec('zstd').decompress(buf, decompressed_size=size, asbytes=True)


def _compress(buf, comp):
    if comp is None:
        return buf
    cid = comp.get('id')
    if cid == 'zlib':
        return zlib.compress(buf, comp.get('level', 1))
    if cid == 'gzip':
        co = zlib.compressobj(comp.get('level', 1), zlib.DEFLATED, 16 + zlib.MAX_WBITS)
        return co.compress(buf) + co.flush()
    raise RuntimeError("ZarrLite cannot write compressor %r" % cid)


class ZarrLiteArray:
    """one zarr v2 array directory: numpy-style basic indexing (ints / slices /
    Ellipsis), reads only the chunks a request touches -- the blockwise driver
    loads one block + halo at a time (stitch_patch_graph.py:443-513)."""

    def __init__(self, path):
        self.path = path
        with open(os.path.join(path, '.zarray')) as f:
            meta = json.load(f)
        assert meta.get('zarr_format', 2) == 2, "zarr v2 only"
        assert meta.get('order', 'C') == 'C', "C order only"
        if meta.get('filters'):
            raise RuntimeError("%s: zarr filters are not supported by ZarrLite" % path)
        self.shape = tuple(int(s) for s in meta['shape'])
        self.chunks = tuple(int(c) for c in meta['chunks'])
        self.dtype = np.dtype(meta['dtype'])
        self.compressor = meta.get('compressor')
        self.fill_value = meta.get('fill_value') or 0
        self.sep = meta.get('dimension_separator', '.')
        self.ndim = len(self.shape)
        self.attrs = {}
        ap = os.path.join(path, '.zattrs')
        if os.path.exists(ap):
            with open(ap) as f:
                self.attrs = json.load(f)

    def __len__(self):
        return self.shape[0]

    def _chunk(self, idx):
        fn = os.path.join(self.path, self.sep.join(str(i) for i in idx))
        if not os.path.exists(fn):
            return np.full(self.chunks, self.fill_value, self.dtype)
        with open(fn, 'rb') as f:
            raw = _decompress(f.read(), self.compressor, fn)
        return np.frombuffer(raw, self.dtype).reshape(self.chunks)

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if Ellipsis in key:
            i = key.index(Ellipsis)
            key = key[:i] + (slice(None),) * (self.ndim - len(key) + 1) + key[i + 1:]
        key = key + (slice(None),) * (self.ndim - len(key))
        sl, squeeze = [], []
        for d, k in enumerate(key):
            if isinstance(k, (int, np.integer)):
                k = int(k) + (self.shape[d] if k < 0 else 0)
                sl.append((k, k + 1))
                squeeze.append(d)
            else:
                a, b, st = k.indices(self.shape[d])
                assert st == 1, "unit strides only"
                sl.append((a, max(a, b)))
        out = np.empty([b - a for a, b in sl], self.dtype)
        rng = [range(a // c, (b - 1) // c + 1) if b > a else range(0)
               for (a, b), c in zip(sl, self.chunks)]
        for idx in itertools.product(*rng):
            ch = self._chunk(idx)
            src, dst = [], []
            for d, i in enumerate(idx):
                c0 = i * self.chunks[d]
                a, b = max(sl[d][0], c0), min(sl[d][1], c0 + self.chunks[d], self.shape[d])
                src.append(slice(a - c0, b - c0))
                dst.append(slice(a - sl[d][0], b - sl[d][0]))
            out[tuple(dst)] = ch[tuple(src)]
        return out.squeeze(axis=tuple(squeeze)) if squeeze else out

    def __array__(self, dtype=None, copy=None):
        a = self[...]
        return a.astype(dtype) if dtype is not None else a


class ZarrLiteGroup:
    def __init__(self, path, mode='r'):
        self.path = os.path.abspath(path)
        self.mode = mode
        if mode in ('w', 'a') and not os.path.exists(os.path.join(self.path, '.zgroup')):
            os.makedirs(self.path, exist_ok=True)
            with open(os.path.join(self.path, '.zgroup'), 'w') as f:
                json.dump({'zarr_format': 2}, f)
        assert os.path.isdir(self.path), self.path + " is not a zarr directory store"

    def _p(self, key):
        return os.path.join(self.path, *key.strip('/').split('/'))

    def __contains__(self, key):
        p = self._p(key)
        return os.path.exists(os.path.join(p, '.zarray')) or \
            os.path.exists(os.path.join(p, '.zgroup'))

    def __getitem__(self, key):
        p = self._p(key)
        if os.path.exists(os.path.join(p, '.zarray')):
            return ZarrLiteArray(p)
        if os.path.isdir(p):
            return ZarrLiteGroup(p, self.mode if self.mode != 'w' else 'a')
        raise KeyError(key)

    def keys(self):
        return [n for n in sorted(os.listdir(self.path)) if not n.startswith('.')]

    def create_dataset(self, name, data=None, shape=None, dtype=None, chunks=None,
                       compressor='default', overwrite=False, **_):
        assert self.mode in ('w', 'a'), "store opened read-only"
        data = np.ascontiguousarray(data, dtype=dtype)
        p = self._p(name)
        if os.path.exists(os.path.join(p, '.zarray')) and not overwrite:
            raise ValueError("array %s exists" % name)
        parts = name.strip('/').split('/')
        for i in range(1, len(parts)):                       # parent groups
            gp = os.path.join(self.path, *parts[:i])
            os.makedirs(gp, exist_ok=True)
            if not os.path.exists(os.path.join(gp, '.zgroup')):
                with open(os.path.join(gp, '.zgroup'), 'w') as f:
                    json.dump({'zarr_format': 2}, f)
        os.makedirs(p, exist_ok=True)
        comp = {'id': 'zlib', 'level': 1} if compressor == 'default' else compressor
        if comp is not None and not isinstance(comp, dict):
            comp = {'id': 'zlib', 'level': 1}                # a numcodecs object: own choice
        chunks = tuple(int(c) for c in (chunks or [max(1, s) for s in data.shape]))
        meta = dict(zarr_format=2, shape=list(data.shape), chunks=list(chunks),
                    dtype=data.dtype.str, compressor=comp, fill_value=0, order='C',
                    filters=None)
        with open(os.path.join(p, '.zarray'), 'w') as f:
            json.dump(meta, f)
        rng = [range(-(-s // c)) for s, c in zip(data.shape, chunks)]
        for idx in itertools.product(*rng):
            ch = np.zeros(chunks, data.dtype)
            src = tuple(slice(i * c, min((i + 1) * c, s))
                        for i, c, s in zip(idx, chunks, data.shape))
            ch[tuple(slice(0, s.stop - s.start) for s in src)] = data[src]
            with open(os.path.join(p, '.'.join(str(i) for i in idx) or '0'), 'wb') as f:
                f.write(_compress(ch.tobytes(), comp))
        return ZarrLiteArray(p)

    def __setitem__(self, name, data):
        self.create_dataset(name, data=data, overwrite=True)


class _NpzContainer:
    """read-only dict view of an .npz ('/' in dataset names kept as is)."""

    def __init__(self, path):
        self.d = np.load(path)

    def __getitem__(self, k):
        return self.d[k]

    def __contains__(self, k):
        return k in self.d.files

    def keys(self):
        return list(self.d.files)


def open_zarr(path, mode='r'):
    try:
        import zarr
    except ImportError:
        return ZarrLiteGroup(path, mode)
    return zarr.open(path, mode)


def open_container(path, mode='r'):
    """zarr directory store, `.hdf` (needs h5py) or `.npz`."""
    if path.endswith('.zarr') or os.path.isdir(path):
        return open_zarr(path, mode)
    if path.endswith('.hdf') or path.endswith('.h5'):
        try:
            import h5py
        except ImportError as e:
            raise RuntimeError("reading %s needs the h5py package" % path) from e
        return h5py.File(path, mode)
    if path.endswith('.npz'):
        return _NpzContainer(path)
    raise RuntimeError("unsupported container " + path)


def write_result(path_noext, datasets, output_format='hdf'):
    """`<sample>.hdf` like the reference when h5py is there; otherwise the same
    datasets go to `<sample>.npz` and a warning says so (downstream evaluation of
    the reference expects .hdf)."""
    if output_format == 'hdf':
        try:
            import h5py
        except ImportError:
            logger.warning("h5py is not installed: writing %s.npz instead of .hdf", path_noext)
            output_format = 'npz'
        else:
            with h5py.File(path_noext + '.hdf', 'w') as f:
                for k, v in datasets.items():
                    f.create_dataset(k, data=v, compression='gzip')
                    f[k].attrs['offset'] = (0, 0, 0)
                    f[k].attrs['resolution'] = (1, 1, 1)
            return path_noext + '.hdf'
    if output_format == 'zarr':
        g = open_zarr(path_noext + '.zarr', 'a')
        for k, v in datasets.items():
            g.create_dataset(k, data=v, overwrite=True)
        return path_noext + '.zarr'
    np.savez_compressed(path_noext + '.npz', **datasets)
    return path_noext + '.npz'
