"""Minimal GeoTIFF reader / writer (numpy + zlib only) for the tile and mask files of the path.

The reference reads tiles with rasterio / GDAL (`data.py:18-28` `open_npy`, `utils.py:40-55` `load_gdal` / `get_y`,
`predict.py:206-215` geotransform + projection of every tile) and writes predictions with `store_tif`
(`predict.py:19-52`: GTiff driver, Byte or Float32, geotransform, projection, optional nodata, `class_zero` un-shift).
Neither library exists in this image, and the hot path only needs the baseline subset of the format, so it is restated
here: classic and BigTIFF, little/big endian, strips or tiles, chunky or planar samples, uncompressed / Deflate / LZW /
PackBits, horizontal predictor, 8/16/32/64-bit integer and float samples.  Georeferencing travels as the raw GeoTIFF
tags (ModelPixelScale / ModelTiepoint / ModelTransformation / GeoKeyDirectory / GeoDoubleParams / GeoAsciiParams):
copying the GeoKey tags from the input tile to the prediction preserves the CRS exactly, which is what
`out_ds.SetProjection(geo_proj)` does in the reference without needing a WKT parser.
"""
from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Optional, Sequence, Tuple, Union

import numpy as np

# tag ids
_W, _H, _BITS, _COMP, _PHOTO, _STRIP_OFF, _SPP, _RPS, _STRIP_CNT = 256, 257, 258, 259, 262, 273, 277, 278, 279
_PLANAR, _PREDICTOR, _TILE_W, _TILE_H, _TILE_OFF, _TILE_CNT, _EXTRA, _SFMT = 284, 317, 322, 323, 324, 325, 338, 339
_PIXSCALE, _TIEPOINT, _TRANSFORM, _GEOKEYS, _GEODOUBLE, _GEOASCII, _GDAL_NODATA = 33550, 33922, 34264, 34735, 34736, 34737, 42113

_TYPE_FMT = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 7: "B", 8: "h", 9: "i", 10: "ii", 11: "f", 12: "d",
             16: "Q", 17: "q", 18: "Q"}
_TYPE_SIZE = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 16: 8, 17: 8, 18: 8}


@dataclass
class GeoInfo:
    """Georeferencing of a raster: GDAL-order geotransform (ulx, xres, xskew, uly, yskew, yres) plus the raw GeoKey tags."""
    geotransform: Tuple[float, float, float, float, float, float] = (0.0, 1.0, 0.0, 0.0, 0.0, 1.0)
    geokeys: Optional[Tuple[int, ...]] = None          # GeoKeyDirectoryTag (SHORT[])
    geodoubles: Optional[Tuple[float, ...]] = None     # GeoDoubleParamsTag
    geoascii: Optional[str] = None                     # GeoAsciiParamsTag
    nodata: Optional[float] = None
    georeferenced: bool = False

    def same_projection(self, other: "GeoInfo") -> bool:
        """what predict.py:209-212 compares with GetProjection() strings"""
        return (self.geokeys, self.geodoubles, self.geoascii) == (other.geokeys, other.geodoubles, other.geoascii)

    def window(self, x: int, y: int) -> "GeoInfo":
        """georeferencing of the sub-window whose upper-left pixel is (x, y) (create_tiles_unet.py:224-226, with the y
        origin computed from the y pixel size - the reference uses the x size there, which is only right for square pixels)"""
        g = self.geotransform
        gt = (g[0] + x * g[1] + y * g[2], g[1], g[2], g[3] + x * g[4] + y * g[5], g[4], g[5])
        return GeoInfo(gt, self.geokeys, self.geodoubles, self.geoascii, self.nodata, self.georeferenced)


class TiffError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------------ decompression
def _lzw_decode(data: bytes) -> bytes:
    """TIFF LZW (MSB-first codes, 9..12 bits, ClearCode 256, EOI 257, 'early change')."""
    out = bytearray()
    table = [bytes([i]) for i in range(256)] + [b"", b""]
    nbits, bitbuf, bitcnt, prev = 9, 0, 0, None
    for byte in data:
        bitbuf = (bitbuf << 8) | byte
        bitcnt += 8
        while bitcnt >= nbits:
            code = (bitbuf >> (bitcnt - nbits)) & ((1 << nbits) - 1)
            bitcnt -= nbits
            if code == 256:
                table = table[:258]
                nbits, prev = 9, None
                continue
            if code == 257:
                return bytes(out)
            if prev is None:
                entry = table[code]
            elif code < len(table):
                entry = table[code]
                table.append(prev + entry[:1])
            else:
                entry = prev + prev[:1]
                table.append(entry)
            out += entry
            prev = entry
            if len(table) >= (1 << nbits) - 1 and nbits < 12:
                nbits += 1
    return bytes(out)


def _packbits_decode(data: bytes) -> bytes:
    out, i, n = bytearray(), 0, len(data)
    while i < n:
        h = data[i]
        i += 1
        if h < 128:
            out += data[i:i + h + 1]
            i += h + 1
        elif h > 128:
            out += data[i:i + 1] * (257 - h)
            i += 1
    return bytes(out)


def _decompress(buf: bytes, comp: int) -> bytes:
    if comp == 1:
        return buf
    if comp in (8, 32946):
        return zlib.decompress(buf)
    if comp == 5:
        return _lzw_decode(buf)
    if comp == 32773:
        return _packbits_decode(buf)
    raise TiffError(f"TIFF compression {comp} is not supported (none, LZW, Deflate and PackBits are)")


# ------------------------------------------------------------------------------------------------------------ reading
def _read_ifd(f, bo: str, big: bool):
    """Parses the first IFD; returns {tag: tuple(values)}."""
    if big:
        (off,) = struct.unpack(bo + "Q", f.read(8))
        f.seek(off)
        (n,) = struct.unpack(bo + "Q", f.read(8))
        esz, cnt_fmt, inl = 20, "Q", 8
    else:
        (off,) = struct.unpack(bo + "I", f.read(4))
        f.seek(off)
        (n,) = struct.unpack(bo + "H", f.read(2))
        esz, cnt_fmt, inl = 12, "I", 4
    raw = f.read(n * esz)
    tags: Dict[int, tuple] = {}
    for i in range(n):
        e = raw[i * esz:(i + 1) * esz]
        tag, typ = struct.unpack(bo + "HH", e[:4])
        (cnt,) = struct.unpack(bo + cnt_fmt, e[4:4 + inl])
        if typ not in _TYPE_SIZE:
            continue
        nbytes = _TYPE_SIZE[typ] * cnt
        val = e[4 + inl:4 + 2 * inl]
        if nbytes > inl:
            (voff,) = struct.unpack(bo + cnt_fmt, val)
            f.seek(voff)
            val = f.read(nbytes)
        else:
            val = val[:nbytes]
        if typ == 2:
            tags[tag] = (val.split(b"\x00")[0].decode("latin-1"),)
        elif typ in (5, 10):
            v = struct.unpack(bo + _TYPE_FMT[typ][0] * (2 * cnt), val)
            tags[tag] = tuple(v[2 * j] / v[2 * j + 1] if v[2 * j + 1] else 0.0 for j in range(cnt))
        else:
            tags[tag] = struct.unpack(bo + _TYPE_FMT[typ] * cnt, val)
    return tags


def _dtype_of(bits: int, sfmt: int, bo: str) -> np.dtype:
    kind = {1: "u", 2: "i", 3: "f"}.get(sfmt)
    if kind is None or bits not in (8, 16, 32, 64) or (kind == "f" and bits < 32):
        raise TiffError(f"unsupported sample format: {bits} bits, SampleFormat {sfmt}")
    return np.dtype(("<" if bo == "<" else ">") + kind + str(bits // 8))


def _geo_from_tags(tags) -> GeoInfo:
    g = GeoInfo()
    if _TRANSFORM in tags:
        m = tags[_TRANSFORM]
        g.geotransform = (m[3], m[0], m[1], m[7], m[4], m[5])
        g.georeferenced = True
    elif _PIXSCALE in tags and _TIEPOINT in tags:
        sx, sy = tags[_PIXSCALE][0], tags[_PIXSCALE][1]
        i, j, _, x, y, _ = tags[_TIEPOINT][:6]
        g.geotransform = (x - i * sx, sx, 0.0, y + j * sy, 0.0, -sy)
        g.georeferenced = True
    if _GEOKEYS in tags:
        g.geokeys = tuple(int(v) for v in tags[_GEOKEYS])
    if _GEODOUBLE in tags:
        g.geodoubles = tuple(float(v) for v in tags[_GEODOUBLE])
    if _GEOASCII in tags:
        g.geoascii = tags[_GEOASCII][0]
    if _GDAL_NODATA in tags:
        try:
            g.nodata = float(tags[_GDAL_NODATA][0])
        except ValueError:
            g.nodata = None
    return g


def read_geotiff(path: Union[str, Path], window: Optional[Tuple[int, int, int, int]] = None):
    """Reads the first image of a (Geo)TIFF as `[bands, H, W]` in the file's sample type (what `rasterio.open(fn).read()`
    returns, data.py:20) together with its `GeoInfo`.  `window=(x, y, w, h)` reads a sub-rectangle (only the strips /
    tiles it touches are decoded), which is how a 20000 x 20000 raster is streamed tile-row by tile-row."""
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(path)
    with open(path, "rb") as f:
        head = f.read(4)
        if head[:2] == b"II":
            bo = "<"
        elif head[:2] == b"MM":
            bo = ">"
        else:
            raise TiffError(f"{path}: not a TIFF file")
        (magic,) = struct.unpack(bo + "H", head[2:4])
        if magic == 43:
            f.read(4)  # offset size (8) and padding
            big = True
        elif magic == 42:
            big = False
        else:
            raise TiffError(f"{path}: bad TIFF magic {magic}")
        tags = _read_ifd(f, bo, big)
        W, H = int(tags[_W][0]), int(tags[_H][0])
        spp = int(tags.get(_SPP, (1,))[0])
        bits = tags.get(_BITS, (1,))
        sfmt = tags.get(_SFMT, (1,))
        if len(set(bits)) != 1 or len(set(sfmt)) != 1:
            raise TiffError("bands with different sample types are not supported")
        dt = _dtype_of(int(bits[0]), int(sfmt[0]), bo)
        comp = int(tags.get(_COMP, (1,))[0])
        planar = int(tags.get(_PLANAR, (1,))[0])
        pred = int(tags.get(_PREDICTOR, (1,))[0])
        if pred not in (1, 2):
            raise TiffError(f"predictor {pred} is not supported")
        x0, y0, ww, wh = window if window is not None else (0, 0, W, H)
        if x0 < 0 or y0 < 0 or x0 + ww > W or y0 + wh > H or ww <= 0 or wh <= 0:
            raise TiffError(f"window {window} outside the {W}x{H} raster")
        out = np.zeros((spp, wh, ww), dtype=dt.newbyteorder("="))
        tiled = _TILE_W in tags
        if tiled:
            bw, bh = int(tags[_TILE_W][0]), int(tags[_TILE_H][0])
            offs, cnts = tags[_TILE_OFF], tags[_TILE_CNT]
        else:
            bw, bh = W, int(tags.get(_RPS, (H,))[0])
            bh = min(bh, H)
            offs, cnts = tags[_STRIP_OFF], tags[_STRIP_CNT]
        nbx, nby = (W + bw - 1) // bw, (H + bh - 1) // bh
        planes = spp if planar == 2 else 1
        chan = 1 if planar == 2 else spp
        for pl in range(planes):
            for by in range(y0 // bh, (y0 + wh - 1) // bh + 1):
                for bx in range(x0 // bw, (x0 + ww - 1) // bw + 1):
                    k = (pl * nby + by) * nbx + bx
                    f.seek(offs[k])
                    buf = _decompress(f.read(cnts[k]), comp)
                    rows = bh if tiled else min(bh, H - by * bh)
                    need = rows * bw * chan * dt.itemsize
                    if len(buf) < need:
                        raise TiffError(f"{path}: block {k} is truncated ({len(buf)} < {need} bytes)")
                    blk = np.frombuffer(buf, dtype=dt, count=rows * bw * chan).reshape(rows, bw, chan)
                    if pred == 2:
                        blk = np.cumsum(blk.astype(dt.newbyteorder("=")), axis=1, dtype=dt.newbyteorder("="))
                    ys, ye = max(y0, by * bh), min(y0 + wh, by * bh + rows)
                    xs, xe = max(x0, bx * bw), min(x0 + ww, (bx + 1) * bw, W)
                    if ys >= ye or xs >= xe:
                        continue
                    sub = blk[ys - by * bh:ye - by * bh, xs - bx * bw:xe - bx * bw, :]
                    if planar == 2:
                        out[pl, ys - y0:ye - y0, xs - x0:xe - x0] = sub[:, :, 0]
                    else:
                        out[:, ys - y0:ye - y0, xs - x0:xe - x0] = np.moveaxis(sub, 2, 0)
        geo = _geo_from_tags(tags)
        if window is not None:
            geo = geo.window(x0, y0)
    return out, geo


def geotiff_info(path: Union[str, Path]) -> Tuple[int, int, int, np.dtype, GeoInfo]:
    """(bands, height, width, dtype, GeoInfo) without decoding pixel data (RasterCount / RasterYSize / RasterXSize /
    GetGeoTransform of the reference's gdal.Open calls)."""
    with open(path, "rb") as f:
        head = f.read(4)
        bo = "<" if head[:2] == b"II" else ">"
        (magic,) = struct.unpack(bo + "H", head[2:4])
        big = magic == 43
        if big:
            f.read(4)
        tags = _read_ifd(f, bo, big)
    dt = _dtype_of(int(tags.get(_BITS, (1,))[0]), int(tags.get(_SFMT, (1,))[0]), bo).newbyteorder("=")
    return int(tags.get(_SPP, (1,))[0]), int(tags[_H][0]), int(tags[_W][0]), dt, _geo_from_tags(tags)


# ------------------------------------------------------------------------------------------------------------ writing
def write_geotiff(path: Union[str, Path], array: np.ndarray, geo: Optional[GeoInfo] = None,
                  nodata: Optional[float] = None, class_zero: bool = False, compress: bool = False,
                  tile: Optional[int] = None, planar: bool = False) -> None:
    """`store_tif` (predict.py:19-52): writes `[bands, H, W]` or `[H, W]` as a striped, pixel-interleaved GeoTIFF (GDAL's
    GTiff default) with the geotransform / GeoKeys of `geo`.  `class_zero`: label 0 becomes `nodata` and every other
    label is decremented, exactly as predict.py:34-36 (`np.where(a == 0, nodata, a - 1)`; with nodata None numpy yields
    an object array there - here 0 stays 0 in that case).  BigTIFF is chosen automatically beyond 4 GB.
    `tile` (a multiple of 16) writes square tiles instead of strips (GDAL `TILED=YES`), `planar` band-sequential samples
    (GDAL `INTERLEAVE=BAND`): the layouts large source rasters usually come in."""
    a = np.asarray(array)
    if a.ndim == 2:
        a = a[None]
    if a.ndim != 3:
        raise ValueError("write_geotiff expects [bands, H, W] or [H, W]")
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if a.dtype == np.int64:      # argmax output of the reference is int64, written through GDT_Byte (predict.py:337-340)
        a = a.astype(np.uint8)
    if class_zero:
        nd = nodata if nodata is not None else 0
        a = np.where(a == 0, np.asarray(nd, dtype=a.dtype), a - np.asarray(1, dtype=a.dtype)).astype(a.dtype)
    kind = a.dtype.kind
    if kind not in "uif":
        raise ValueError(f"dtype {a.dtype} cannot be stored")
    sfmt = {"u": 1, "i": 2, "f": 3}[kind]
    bands, H, W = a.shape
    data = np.ascontiguousarray(np.moveaxis(a, 0, 2)).astype(a.dtype.newbyteorder("<"), copy=False)
    row_bytes = W * bands * a.dtype.itemsize
    rps = max(1, min(H, (1 << 20) // max(1, row_bytes)))
    if tile is not None and (tile <= 0 or tile % 16):
        raise ValueError("tile size must be a positive multiple of 16")
    planes = [data[:, :, b:b + 1] for b in range(bands)] if planar else [data]
    strips = []
    for pl in planes:
        if tile is None:
            for r in range(0, H, rps):
                raw = np.ascontiguousarray(pl[r:r + rps]).tobytes()
                strips.append(zlib.compress(raw, 6) if compress else raw)
        else:
            for ty in range(0, H, tile):
                for tx in range(0, W, tile):
                    blk = np.zeros((tile, tile, pl.shape[2]), dtype=pl.dtype)      # edge tiles are padded to full size
                    sub = pl[ty:ty + tile, tx:tx + tile]
                    blk[:sub.shape[0], :sub.shape[1]] = sub
                    raw = blk.tobytes()
                    strips.append(zlib.compress(raw, 6) if compress else raw)
    total = sum(len(s) for s in strips)
    big = total + 4096 + 16 * len(strips) > 0xFFFF0000
    geo = geo or GeoInfo()

    entries = []   # (tag, type, count, packed bytes)

    def add(tag, typ, values):
        if typ == 2:
            b = values.encode("latin-1") + b"\x00"
            entries.append((tag, 2, len(b), b))
        else:
            fmt = _TYPE_FMT[typ]
            entries.append((tag, typ, len(values), struct.pack("<" + fmt * len(values), *values)))

    off_t = 16 if big else 4
    add(_W, 4, [W]); add(_H, 4, [H])
    add(_BITS, 3, [a.dtype.itemsize * 8] * bands)
    add(_COMP, 3, [8 if compress else 1])
    rgb = bands >= 3 and a.dtype == np.uint8        # GDAL's GTiff default: RGB photometric for >= 3 byte bands
    add(_PHOTO, 3, [2 if rgb else 1])
    off_tag, cnt_tag = (_STRIP_OFF, _STRIP_CNT) if tile is None else (_TILE_OFF, _TILE_CNT)
    add(off_tag, off_t, [0] * len(strips))               # patched below
    add(_SPP, 3, [bands])
    if tile is None:
        add(_RPS, 4, [rps])
    else:
        add(_TILE_W, 4, [tile]); add(_TILE_H, 4, [tile])
    add(cnt_tag, off_t, [len(s) for s in strips])
    add(_PLANAR, 3, [2 if planar else 1])
    n_extra = bands - 3 if rgb else bands - 1
    if n_extra > 0:
        add(_EXTRA, 3, [0] * n_extra)
    add(_SFMT, 3, [sfmt] * bands)
    if geo.georeferenced:
        g = geo.geotransform
        if g[2] == 0.0 and g[4] == 0.0:
            add(_PIXSCALE, 12, [g[1], -g[5], 0.0])
            add(_TIEPOINT, 12, [0.0, 0.0, 0.0, g[0], g[3], 0.0])
        else:
            add(_TRANSFORM, 12, [g[1], g[2], 0.0, g[0], g[4], g[5], 0.0, g[3], 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0])
    if geo.geokeys:
        add(_GEOKEYS, 3, list(geo.geokeys))
    if geo.geodoubles:
        add(_GEODOUBLE, 12, list(geo.geodoubles))
    if geo.geoascii:
        add(_GEOASCII, 2, geo.geoascii)
    nd = nodata if nodata is not None else None
    if nd is not None:
        add(_GDAL_NODATA, 2, repr(int(nd)) if float(nd).is_integer() else repr(float(nd)))
    entries.sort(key=lambda e: e[0])

    head = 16 if big else 8
    esz, inl = (20, 8) if big else (12, 4)
    ifd_size = (8 if big else 2) + len(entries) * esz + (8 if big else 4)
    ext_off = head + ifd_size
    ext = bytearray()
    placed = []
    for tag, typ, cnt, b in entries:
        if len(b) <= inl:
            placed.append((tag, typ, cnt, b.ljust(inl, b"\x00"), None))
        else:
            if len(ext) % 2:
                ext += b"\x00"
            placed.append((tag, typ, cnt, None, ext_off + len(ext)))
            ext += b
    data_off = ext_off + len(ext)
    data_off += (-data_off) % 16
    # strip offsets are known now: rewrite that entry's payload
    offs, o = [], data_off
    for s in strips:
        offs.append(o)
        o += len(s)
    off_bytes = struct.pack("<" + _TYPE_FMT[off_t] * len(offs), *offs)
    with open(path, "wb") as f:
        if big:
            f.write(struct.pack("<2sHHHQ", b"II", 43, 8, 0, head))
            f.write(struct.pack("<Q", len(entries)))
        else:
            f.write(struct.pack("<2sHI", b"II", 42, head))
            f.write(struct.pack("<H", len(entries)))
        ext = bytearray(ext)
        for tag, typ, cnt, inline, where in placed:
            if tag == off_tag:
                if where is None:
                    inline = off_bytes.ljust(inl, b"\x00")
                else:
                    ext[where - ext_off:where - ext_off + len(off_bytes)] = off_bytes
            f.write(struct.pack("<HH", tag, typ))
            f.write(struct.pack("<Q" if big else "<I", cnt))
            f.write(inline if where is None else struct.pack("<Q" if big else "<I", where))
        f.write(struct.pack("<Q" if big else "<I", 0))
        f.write(bytes(ext))
        f.write(b"\x00" * (data_off - ext_off - len(ext)))
        for s in strips:
            f.write(s)


# ------------------------------------------------------------------------------------------ reference-shaped helpers
def open_tile(fn: Union[str, Path], chnls: Optional[Sequence[int]] = None) -> np.ndarray:
    """`open_npy` (data.py:18-28): `[bands, H, W]` of an image tile.  uint8 / uint16 / int16 tiles keep their raw band
    values (the device kernels apply the int32 -> float32 cast of data.py:24 and the divisions by 255 of the reference's
    batch transforms, `unet_b200.network.input_contract`); any other sample type is returned as float32 divided by 255."""
    a, _ = read_geotiff(fn)
    if chnls is not None:
        a = a[list(chnls)]
    if a.dtype in (np.uint8, np.uint16, np.int16):
        return a
    return a.astype(np.int32).astype(np.float32) / 255.0 if a.dtype.kind in "ui" else a.astype(np.float32) / 255.0


def open_mask(fn: Union[str, Path]) -> np.ndarray:
    """`get_y` (utils.py:51-55): band 1 of the mask tile next to an image tile (`img_tiles` -> `mask_tiles`)."""
    a, _ = read_geotiff(str(fn).replace("img_tiles", "mask_tiles"))
    return a[0]
