"""Point-cloud file IO of the reference API: the native PCRP SoA binary and CSV.

File formats follow src/io/point_cloud_io.cpp of the reference:
  PCRP  magic "PCRP" u32 | version 1 u32 | num_points u64 | num_channels u32 | crs_wkt_len u32 | wkt |
        per channel {name_len u16, name, dtype u8} | x f64[n] | y f64[n] | channel arrays in table order
        (write_pcr_binary :74-148, read_pcr_binary_info :151-216)
  CSV   header "x,y,<channel>...", one row per point, channels read back as Float64 (:293-470)

PointCloudReader streams PCRP in chunks by SEEKING into each SoA array; the reference's
read_chunk_pcr (:574-612) reads x, y and the channels back to back from the current position, which
is only right when one chunk covers the whole file — that defect is not reproduced.  LAS/LAZ are
NotImplemented upstream; uncompressed LAS is read (and written, format 0) here through _las.py, LAZ
stays NotImplemented.  Host-side numpy code: file IO sits in front of the hot path.
"""
import os
import struct

import numpy as np

from . import _las

MAGIC = 0x50524350


def _api():
    from . import (PointCloud, PointCloudInfo, ChannelDesc, DataType, CRS, BBox, PointCloudFormat,
                   MemoryLocation, _NP_DTYPE)
    return PointCloud, PointCloudInfo, ChannelDesc, DataType, CRS, BBox, PointCloudFormat, MemoryLocation, _NP_DTYPE


def _detect(path, fmt):
    PointCloudFormat = _api()[6]
    fmt = PointCloudFormat(int(fmt))
    if fmt != PointCloudFormat.Auto:
        return fmt
    low = str(path).lower()
    for ext, f in ((".pcr", PointCloudFormat.PCR_Binary), (".pcrp", PointCloudFormat.PCR_Binary),
                   (".csv", PointCloudFormat.CSV), (".las", PointCloudFormat.LAS), (".laz", PointCloudFormat.LAZ)):
        if low.endswith(ext):
            return f
    try:
        with open(path, "rb") as f:
            if struct.unpack("<I", f.read(4))[0] == MAGIC:
                return PointCloudFormat.PCR_Binary
    except Exception:
        pass
    return PointCloudFormat.CSV


def _read_pcrp_header(f):
    """-> (num_points, [(name, DataType)], wkt, body_offset)"""
    DataType = _api()[3]
    head = f.read(20)
    if len(head) < 20:
        raise RuntimeError("failed to read header")
    magic, version, n, nch = struct.unpack("<IIQI", head)
    if magic != MAGIC:
        raise RuntimeError("invalid magic number (not a PCRP file)")
    if version != 1:
        raise RuntimeError(f"unsupported version {version}")
    (wlen,) = struct.unpack("<I", f.read(4))
    wkt = f.read(wlen).decode("utf-8", "replace") if wlen else ""
    chans = []
    for _ in range(nch):
        (nl,) = struct.unpack("<H", f.read(2))
        name = f.read(nl).decode("utf-8", "replace")
        (dt,) = struct.unpack("<B", f.read(1))
        chans.append((name, DataType(dt)))
    return n, chans, wkt, f.tell()


def read_point_cloud_info(path, fmt=4):
    PointCloud, PointCloudInfo, ChannelDesc, DataType, CRS, BBox, PointCloudFormat, _, _ = _api()
    fmt = _detect(path, fmt)
    info = PointCloudInfo()
    if fmt == PointCloudFormat.PCR_Binary:
        try:
            with open(path, "rb") as f:
                n, chans, wkt, _ = _read_pcrp_header(f)
        except OSError:
            raise RuntimeError(f"failed to open file: {path}")
        info.num_points = n
        info.channels = [ChannelDesc(nm, dt) for nm, dt in chans]
        info.crs = CRS.from_wkt(wkt) if wkt else CRS()
    elif fmt == PointCloudFormat.CSV:
        try:
            with open(path) as f:
                header = f.readline().rstrip("\n").split(",")
                n = sum(1 for line in f if line.strip())
        except OSError:
            raise RuntimeError(f"failed to open file: {path}")
        if len(header) < 2 or header[0] != "x" or header[1] != "y":
            raise RuntimeError("CSV must start with x,y columns")
        info.num_points = n
        info.channels = [ChannelDesc(nm, DataType.Float64) for nm in header[2:]]
    elif fmt == PointCloudFormat.LAS:
        try:
            with open(path, "rb") as f:
                h = _las.read_header(f)
        except OSError:
            raise RuntimeError(f"failed to open file: {path}")
        if h.compressed:
            raise RuntimeError("LAZ (compressed LAS) format not yet implemented")
        info.num_points = int(h.num_points)
        names = list(_las.CHANNELS) + (["gps_time"] if "gps_time" in _las.record_dtype(h).names else [])
        info.channels = [ChannelDesc(nm, DataType.Float32) for nm in names]
        if h.epsg:
            info.crs = CRS.from_epsg(h.epsg)
        info.bounds = BBox()
        if info.num_points:
            info.bounds.min_x, info.bounds.min_y, info.bounds.max_x, info.bounds.max_y = h.min_x, h.min_y, h.max_x, h.max_y
        return info
    else:
        raise RuntimeError("LAS/LAZ format not yet implemented")
    info.bounds = BBox()          # not stored in either format
    return info


def write_point_cloud(path, cloud, fmt=0):
    PointCloud, _, _, DataType, _, _, PointCloudFormat, MemoryLocation, NP = _api()
    fmt = PointCloudFormat(int(fmt))
    if cloud.location() == MemoryLocation.Device:
        raise RuntimeError("cloud must be on host")
    n = cloud.count()
    names = cloud.channel_names()
    if fmt == PointCloudFormat.PCR_Binary:
        try:
            f = open(path, "wb")
        except OSError:
            raise RuntimeError(f"failed to open file for writing: {path}")
        with f:
            wkt = cloud.crs().wkt.encode()
            f.write(struct.pack("<IIQI", MAGIC, 1, n, len(names)))
            f.write(struct.pack("<I", len(wkt)) + wkt)
            for nm in names:
                b = nm.encode()
                f.write(struct.pack("<H", len(b)) + b + struct.pack("<B", int(cloud.channel(nm).dtype)))
            f.write(np.ascontiguousarray(cloud.x_array(), "<f8").tobytes())
            f.write(np.ascontiguousarray(cloud.y_array(), "<f8").tobytes())
            for nm in names:
                desc, storage = cloud._channels[nm]
                f.write(np.ascontiguousarray(cloud._view(storage, NP[desc.dtype], n)).tobytes())
    elif fmt == PointCloudFormat.CSV:
        cols = [cloud.x_array(), cloud.y_array()]
        for nm in names:
            desc, storage = cloud._channels[nm]
            if desc.dtype not in (DataType.Float32, DataType.Float64, DataType.Int32, DataType.UInt32):
                raise RuntimeError("unsupported channel data type")
            cols.append(cloud._view(storage, NP[desc.dtype], n))
        with open(path, "w") as f:
            f.write(",".join(["x", "y"] + names) + "\n")
            for row in zip(*cols):
                f.write(",".join(f"{v:.15g}" for v in row) + "\n")
    elif fmt == PointCloudFormat.LAS:
        ch = {}
        for nm in names:
            desc, storage = cloud._channels[nm]
            ch[nm] = np.asarray(cloud._view(storage, NP[desc.dtype], n), np.float64)
        _las.write(path, cloud.x_array(), cloud.y_array(), ch, epsg=cloud.crs().epsg)
    else:
        raise RuntimeError("LAS/LAZ format not yet implemented")


def read_point_cloud(path, fmt=4):
    PointCloud, _, _, DataType, CRS, _, PointCloudFormat, _, NP = _api()
    reader = PointCloudReader.open(path, fmt)
    n = reader.info().num_points
    cloud = PointCloud.create(max(n, 1))
    got = reader.read_chunk(cloud, max(n, 1))
    if got != n:
        raise RuntimeError(f"Failed to read point cloud: {path}")
    cloud.set_crs(reader.info().crs)
    return cloud


class PointCloudReader:
    """Streaming reader (PointCloudReader, include/pcr/io/point_cloud_io.h:69-97): open -> info(),
    read_chunk(cloud, max_points) fills `cloud` (resizing it and adding missing channels) and returns
    the number of points read; eof(); rewind().  Chunks can be fed straight to Pipeline.ingest."""

    def __init__(self):
        self._info = None
        self._fmt = None
        self._path = None
        self._pos = 0
        self._body = 0
        self._f = None

    @staticmethod
    def open(path, fmt=4):
        PointCloudFormat = _api()[6]
        r = PointCloudReader()
        r._path = str(path)
        r._fmt = _detect(path, fmt)
        try:
            r._info = read_point_cloud_info(path, r._fmt)
            if r._fmt == PointCloudFormat.PCR_Binary:
                r._f = open(path, "rb")
                _, _, _, r._body = _read_pcrp_header(r._f)
            elif r._fmt == PointCloudFormat.LAS:
                r._f = open(path, "rb")
                r._las = _las.read_header(r._f)
                r._body = r._las.point_offset
            else:
                r._f = open(path)
                r._f.readline()
        except (OSError, RuntimeError) as e:
            raise RuntimeError(f"Failed to open point cloud: {path} ({e})")
        return r

    def info(self):
        return self._info

    def eof(self):
        return self._pos >= self._info.num_points

    def rewind(self):
        PointCloudFormat = _api()[6]
        self._pos = 0
        if self._fmt == PointCloudFormat.CSV:
            self._f.seek(0)
            self._f.readline()

    def read_chunk(self, cloud, max_points):
        _, _, _, DataType, _, _, PointCloudFormat, _, NP = _api()
        n_total = self._info.num_points
        if self._pos >= n_total:
            return 0
        cnt = int(min(max_points, n_total - self._pos, cloud.capacity()))
        for ch in self._info.channels:
            if not cloud.has_channel(ch.name):
                cloud.add_channel(ch.name, ch.dtype)
        if self._fmt == PointCloudFormat.PCR_Binary:
            def grab(array_offset, dtype):
                item = np.dtype(dtype).itemsize
                self._f.seek(self._body + array_offset + self._pos * item)
                return np.frombuffer(self._f.read(cnt * item), dtype=dtype)
            cloud.set_x_array(grab(0, "<f8"))
            cloud.set_y_array(grab(n_total * 8, "<f8"))
            off = 2 * n_total * 8
            for ch in self._info.channels:
                dt = np.dtype(NP[ch.dtype]).newbyteorder("<")
                desc, storage = cloud._channels[ch.name]
                cloud._store(storage, NP[desc.dtype], grab(off, dt), ch.name)
                off += n_total * dt.itemsize
        elif self._fmt == PointCloudFormat.LAS:
            dt = _las.record_dtype(self._las)
            self._f.seek(self._body + self._pos * dt.itemsize)
            recs = np.frombuffer(self._f.read(cnt * dt.itemsize), dtype=dt)
            cnt = len(recs)
            x, y, ch = _las.decode(self._las, recs)
            if "gps_time" in dt.names:
                ch["gps_time"] = recs["gps_time"].astype(np.float32)
            cloud.set_x_array(x)
            cloud.set_y_array(y)
            for nm, arr in ch.items():
                desc, storage = cloud._channels[nm]
                cloud._store(storage, NP[desc.dtype], arr, nm)
        else:
            rows = []
            while len(rows) < cnt:
                line = self._f.readline()
                if not line:
                    break
                if line.strip():
                    rows.append([float(t) for t in line.rstrip("\n").split(",")])
            cnt = len(rows)
            a = np.array(rows, np.float64).reshape(cnt, 2 + len(self._info.channels))
            cloud.set_x_array(a[:, 0])
            cloud.set_y_array(a[:, 1])
            for k, ch in enumerate(self._info.channels):
                desc, storage = cloud._channels[ch.name]
                cloud._store(storage, NP[desc.dtype], a[:, 2 + k], ch.name)
        self._pos += cnt
        return cnt
