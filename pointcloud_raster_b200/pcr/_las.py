"""Uncompressed LAS 1.0-1.4 point files (ASPRS LAS specification; point data record formats 0-10) as a
source for Pipeline.ingest.  The reference declares the format and returns "LAS/LAZ format not yet
implemented" (src/io/point_cloud_io.cpp); its LiDAR script goes through laspy instead
(scripts/data/test_dc_lidar.py:57-100).  Here the public header block and the fixed part of the point
records are decoded with numpy: x, y as f64 (integer * scale + offset) and the Float32 channels the
pipeline can reduce: z, intensity, classification, return_number, number_of_returns (+ gps_time for
the formats that carry it).  LAZ (compressed) stays NotImplemented.  Host-side code in front of the
hot path."""
import struct

import numpy as np

CHANNELS = ("z", "intensity", "classification", "return_number", "number_of_returns")
_GPS_FORMATS = (1, 3, 4, 5, 6, 7, 8, 9, 10)
_MIN_RECORD = {0: 20, 1: 28, 2: 26, 3: 34, 4: 57, 5: 63, 6: 30, 7: 36, 8: 38, 9: 59, 10: 67}


class LasHeader:
    pass


def read_header(f):
    head = f.read(227)
    if len(head) < 227 or head[:4] != b"LASF":
        raise RuntimeError("not a LAS file (missing LASF signature)")
    h = LasHeader()
    h.version = (head[24], head[25])
    h.header_size, h.point_offset, h.num_vlrs = struct.unpack_from("<HII", head, 94)
    h.point_format = head[104] & 0x3F          # bits 6-7 flag LAZ compression in laszip files
    h.compressed = (head[104] & 0xC0) != 0
    (h.record_length,) = struct.unpack_from("<H", head, 105)
    (h.num_points,) = struct.unpack_from("<I", head, 107)
    h.scale = struct.unpack_from("<3d", head, 131)
    h.offset = struct.unpack_from("<3d", head, 155)
    h.max_x, h.min_x, h.max_y, h.min_y, h.max_z, h.min_z = struct.unpack_from("<6d", head, 179)
    if h.version >= (1, 4) and h.header_size >= 375:
        rest = f.read(h.header_size - 227)
        (n64,) = struct.unpack_from("<Q", rest, 247 - 227)
        if n64:
            h.num_points = n64
    if h.point_format not in _MIN_RECORD:
        raise RuntimeError(f"unsupported LAS point data record format {h.point_format}")
    if h.record_length < _MIN_RECORD[h.point_format]:
        raise RuntimeError("LAS point record length is shorter than its format requires")
    h.epsg = _epsg_from_vlrs(f, h)
    return h


def _epsg_from_vlrs(f, h):
    """GeoKeyDirectoryTag VLR (user id LASF_Projection, record id 34735): ProjectedCSTypeGeoKey (3072),
    else GeographicTypeGeoKey (2048)."""
    try:
        f.seek(h.header_size)
        for _ in range(h.num_vlrs):
            vh = f.read(54)
            if len(vh) < 54:
                break
            user = vh[2:18].rstrip(b"\0")
            rec_id, length = struct.unpack_from("<HH", vh, 18)
            body = f.read(length)
            if user == b"LASF_Projection" and rec_id == 34735 and len(body) >= 8:
                nkeys = struct.unpack_from("<H", body, 6)[0]
                keys = {}
                for k in range(nkeys):
                    kid, loc, _cnt, val = struct.unpack_from("<4H", body, 8 + 8 * k)
                    if loc == 0:
                        keys[kid] = val
                for kid in (3072, 2048):
                    if keys.get(kid, 0) not in (0, 32767):
                        return int(keys[kid])
    except (OSError, struct.error):
        pass
    return 0


def record_dtype(h):
    fields = [("X", "<i4"), ("Y", "<i4"), ("Z", "<i4"), ("intensity", "<u2"), ("returns", "u1")]
    if h.point_format >= 6:
        fields += [("flags", "u1"), ("classification", "u1"), ("user_data", "u1"), ("scan_angle", "<i2"),
                   ("source", "<u2"), ("gps_time", "<f8")]
    else:
        fields += [("classification", "u1"), ("scan_angle", "i1"), ("user_data", "u1"), ("source", "<u2")]
        if h.point_format in _GPS_FORMATS:
            fields.append(("gps_time", "<f8"))
    names = [n for n, _ in fields]
    formats = [t for _, t in fields]
    offsets, at = [], 0
    for t in formats:
        offsets.append(at)
        at += np.dtype(t).itemsize
    return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": h.record_length})


def decode(h, recs):
    """structured records -> (x f64, y f64, {channel: f32})"""
    x = recs["X"].astype(np.float64) * h.scale[0] + h.offset[0]
    y = recs["Y"].astype(np.float64) * h.scale[1] + h.offset[1]
    r = recs["returns"]
    if h.point_format >= 6:
        ret, nret = r & 0x0F, r >> 4
        cls = recs["classification"]
    else:
        ret, nret = r & 0x07, (r >> 3) & 0x07
        cls = recs["classification"] & 0x1F            # bits 5-7: synthetic / key-point / withheld
    ch = {"z": (recs["Z"].astype(np.float64) * h.scale[2] + h.offset[2]).astype(np.float32),
          "intensity": recs["intensity"].astype(np.float32),
          "classification": cls.astype(np.float32),
          "return_number": ret.astype(np.float32),
          "number_of_returns": nret.astype(np.float32)}
    return x, y, ch


def write(path, x, y, ch, epsg=0, scale=0.001):
    """LAS 1.2, point data record format 0 (what every LAS reader accepts); coordinates are quantised
    to `scale` around an integral offset, as the format requires."""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    n = len(x)
    z = np.asarray(ch.get("z", np.zeros(n)), np.float64)
    off = [float(np.floor(a.min())) if n else 0.0 for a in (x, y, z)]
    rec = np.zeros(n, np.dtype([("X", "<i4"), ("Y", "<i4"), ("Z", "<i4"), ("intensity", "<u2"), ("returns", "u1"),
                                ("classification", "u1"), ("scan_angle", "i1"), ("user_data", "u1"), ("source", "<u2")]))
    rec["X"] = np.rint((x - off[0]) / scale)
    rec["Y"] = np.rint((y - off[1]) / scale)
    rec["Z"] = np.rint((z - off[2]) / scale)
    rec["intensity"] = np.clip(np.asarray(ch.get("intensity", np.zeros(n))), 0, 65535)
    rec["classification"] = np.asarray(ch.get("classification", np.zeros(n))).astype(np.uint8) & 0x1F
    ret = np.asarray(ch.get("return_number", np.ones(n))).astype(np.uint8) & 7
    nret = np.asarray(ch.get("number_of_returns", np.ones(n))).astype(np.uint8) & 7
    rec["returns"] = ret | (nret << 3)
    vlr = b""
    if epsg:
        body = struct.pack("<4H", 1, 1, 0, 1) + struct.pack("<4H", 3072, 0, 1, int(epsg))
        vlr = struct.pack("<H16sHH32s", 0, b"LASF_Projection", 34735, len(body), b"GeoKeyDirectoryTag") + body
    head = bytearray(227)
    head[0:4] = b"LASF"
    head[24], head[25] = 1, 2
    head[26:58] = b"pcr-b200".ljust(32, b"\0")
    head[58:90] = b"pointcloud_raster_b200".ljust(32, b"\0")
    struct.pack_into("<HII", head, 94, 227, 227 + len(vlr), 1 if vlr else 0)
    head[104] = 0
    struct.pack_into("<H", head, 105, 20)
    struct.pack_into("<I", head, 107, n)
    struct.pack_into("<I", head, 111, n)
    struct.pack_into("<3d", head, 131, scale, scale, scale)
    struct.pack_into("<3d", head, 155, *off)
    if n:
        qx = rec["X"] * scale + off[0]; qy = rec["Y"] * scale + off[1]; qz = rec["Z"] * scale + off[2]
        struct.pack_into("<6d", head, 179, qx.max(), qx.min(), qy.max(), qy.min(), qz.max(), qz.min())
    with open(path, "wb") as f:
        f.write(bytes(head) + vlr + rec.tobytes())
