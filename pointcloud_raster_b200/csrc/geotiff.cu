// geotiff.cu — GDAL-free GeoTIFF writer / header reader (host-only C++).
//
// Stands in for write_geotiff / read_geotiff_info of the reference
// (src/io/grid_io.cpp:39-182,395-445), which are thin GDAL wrappers: Float32,
// one sample per band, tiled (default 256x256), BigTIFF by default, nodata = NaN,
// band descriptions, geotransform from GridConfig::gdal_geotransform
// (src/core/grid_config.cpp:93-110), CRS from the EPSG code when known.
// Layout written here: little-endian (Big)TIFF, PlanarConfiguration = 2 (one plane
// per band, matching the band-major result memory), compression NONE or DEFLATE
// (zlib, tag value 8).  GeoTIFF keys: ModelPixelScale + ModelTiepoint (north-up) or
// ModelTransformation (cell_size_y > 0), GeoKeyDirectory with raster type
// PixelIsArea and the projected / geographic EPSG code; GDAL_NODATA ("nan") and
// GDAL_METADATA (band descriptions) so GDAL-based readers see what the reference
// writes.  Compression NONE, LZW (the reference's default; TIFF 6.0 variable-width codes, MSB first,
// early change) or DEFLATE; ZSTD is rejected with NotImplemented.  cloud_optimized adds the reference's
// overview pyramid (levels 2, 4, ... while min(width, height) / level >= 256, src/io/grid_io.cpp:155-176)
// as reduced-resolution IFDs, NaN-aware 2 x 2 AVERAGE resampling.  TiledGeoTiffWriter (grid_io.h:44-70)
// assembles tiles in host memory and writes the file at close().
#include "../../include/pcr_b200.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <atomic>
#include <vector>

namespace pcrb {
extern thread_local std::string g_geotiff_error;
thread_local std::string g_geotiff_error;
}

namespace {

enum : uint16_t { T_BYTE = 1, T_ASCII = 2, T_SHORT = 3, T_LONG = 4, T_DOUBLE = 12, T_LONG8 = 16 };

struct Entry {
    uint16_t tag, type;
    uint64_t count;
    std::vector<uint8_t> data;   // raw little-endian payload
};

template <typename T>
void put(std::vector<uint8_t>& v, T x)
{
    const uint8_t* p = reinterpret_cast<const uint8_t*>(&x);
    v.insert(v.end(), p, p + sizeof(T));
}

Entry shorts(uint16_t tag, const std::vector<uint16_t>& xs)
{
    Entry e{tag, T_SHORT, xs.size(), {}};
    for (uint16_t x : xs) put(e.data, x);
    return e;
}
Entry longs(uint16_t tag, const std::vector<uint32_t>& xs)
{
    Entry e{tag, T_LONG, xs.size(), {}};
    for (uint32_t x : xs) put(e.data, x);
    return e;
}
Entry doubles(uint16_t tag, const std::vector<double>& xs)
{
    Entry e{tag, T_DOUBLE, xs.size(), {}};
    for (double x : xs) put(e.data, x);
    return e;
}
Entry ascii(uint16_t tag, const std::string& s)
{
    Entry e{tag, T_ASCII, s.size() + 1, {}};
    e.data.assign(s.begin(), s.end());
    e.data.push_back(0);
    return e;
}
Entry offsets(uint16_t tag, const std::vector<uint64_t>& xs, bool big)
{
    Entry e{tag, static_cast<uint16_t>(big ? T_LONG8 : T_LONG), xs.size(), {}};
    for (uint64_t x : xs) { if (big) put(e.data, x); else put(e.data, static_cast<uint32_t>(x)); }
    return e;
}

std::string xml_escape(const std::string& s)
{
    std::string o;
    for (char c : s) {
        if (c == '&') o += "&amp;"; else if (c == '<') o += "&lt;"; else if (c == '>') o += "&gt;";
        else if (c == '"') o += "&quot;"; else o += c;
    }
    return o;
}

int fail(int code, const std::string& msg)
{
    pcrb::g_geotiff_error = msg;
    return code;
}

// ---- TIFF LZW (compression 5): 9..12-bit codes, MSB first, Clear = 256, EOI = 257, "early change" ----
void lzw_encode(const uint8_t* in, size_t n, std::vector<uint8_t>& out)
{
    out.clear();
    uint32_t acc = 0; int nbits = 0;
    auto emit = [&](uint32_t code, int width) {
        acc = (acc << width) | code; nbits += width;
        while (nbits >= 8) { out.push_back(static_cast<uint8_t>(acc >> (nbits - 8))); nbits -= 8; }
    };
    // dictionary: open-addressing hash of (prefix code << 8 | byte) -> code
    constexpr int kHash = 1 << 14;
    std::vector<int32_t> hkey(kHash), hval(kHash);
    auto reset = [&] { std::fill(hkey.begin(), hkey.end(), -1); };
    reset();
    int next = 258, width = 9;
    emit(256, width);
    if (n == 0) { emit(257, width); if (nbits) out.push_back(static_cast<uint8_t>(acc << (8 - nbits))); return; }
    int prefix = in[0];
    for (size_t i = 1; i < n; ++i) {
        const int c = in[i];
        const int32_t key = (prefix << 8) | c;
        uint32_t h = (static_cast<uint32_t>(key) * 2654435761u) >> 18;
        int found = -1;
        while (hkey[h] != -1) { if (hkey[h] == key) { found = hval[h]; break; } h = (h + 1) & (kHash - 1); }
        if (found >= 0) { prefix = found; continue; }
        emit(static_cast<uint32_t>(prefix), width);
        hkey[h] = key; hval[h] = next++;
        // libtiff's rule (tif_lzw.c LZWEncode): the encoder's table runs one entry ahead of the decoder's, so it
        // widens at 512 / 1024 / 2048 where the decoder ("early change") widens at 511 / 1023 / 2047
        if (next == 4094) { emit(256, width); reset(); next = 258; width = 9; }
        else if (next == 512 || next == 1024 || next == 2048) ++width;
        prefix = c;
    }
    emit(static_cast<uint32_t>(prefix), width);
    ++next;                                                                // LZWPostEncode
    if (next == 4094) { emit(256, width); width = 9; }
    else if (next == 512 || next == 1024 || next == 2048) ++width;
    emit(257, width);
    if (nbits) out.push_back(static_cast<uint8_t>(acc << (8 - nbits)));
}

bool lzw_decode(const uint8_t* in, size_t n, uint8_t* out, size_t cap)
{
    std::vector<int32_t> prefix(4096, -1);
    std::vector<uint8_t> suffix(4096, 0), stack;
    for (int i = 0; i < 256; ++i) suffix[i] = static_cast<uint8_t>(i);
    size_t ip = 0, op = 0; uint32_t acc = 0; int nbits = 0, width = 9, next = 258, prev = -1;
    for (;;) {
        while (nbits < width) { if (ip >= n) return op == cap; acc = (acc << 8) | in[ip++]; nbits += 8; }
        const int code = static_cast<int>((acc >> (nbits - width)) & ((1u << width) - 1)); nbits -= width;
        if (code == 257) break;
        if (code == 256) { next = 258; width = 9; prev = -1; continue; }
        int cur = code;
        stack.clear();
        if (prev < 0) { if (code > 255) return false; }
        else if (code >= next) {
            if (code != next) return false;
            int t = prev; while (t > 255) t = prefix[t];
            stack.push_back(static_cast<uint8_t>(t));                      // KwKwK: first byte of the previous string
            cur = prev;
        }
        while (cur > 255) { stack.push_back(suffix[cur]); cur = prefix[cur]; }
        stack.push_back(static_cast<uint8_t>(cur));
        if (op + stack.size() > cap) return false;
        for (size_t k = stack.size(); k-- > 0;) out[op++] = stack[k];
        if (prev >= 0 && next < 4096) { prefix[next] = prev; suffix[next] = static_cast<uint8_t>(cur); ++next; }
        if (next == 511 || next == 1023 || next == 2047) ++width;
        prev = code;
    }
    return op == cap;
}

// one image (the full raster or an overview level): pixel data of all bands already written
struct Image { int W, H; std::vector<uint64_t> tile_off, tile_len; };

// NaN-aware 2 x 2 box average (GDAL "AVERAGE" ignores nodata)
void halve(const std::vector<float>& src, int W, int H, std::vector<float>& dst, int& w2, int& h2)
{
    w2 = (W + 1) / 2; h2 = (H + 1) / 2;
    dst.assign(static_cast<size_t>(w2) * h2, std::numeric_limits<float>::quiet_NaN());
    for (int r = 0; r < h2; ++r)
        for (int c = 0; c < w2; ++c) {
            double sum = 0; int cnt = 0;
            for (int dr = 0; dr < 2; ++dr)
                for (int dc = 0; dc < 2; ++dc) {
                    const int rr = 2 * r + dr, cc = 2 * c + dc;
                    if (rr >= H || cc >= W) continue;
                    const float v = src[static_cast<size_t>(rr) * W + cc];
                    if (v == v) { sum += v; ++cnt; }
                }
            if (cnt) dst[static_cast<size_t>(r) * w2 + c] = static_cast<float>(sum / cnt);
        }
}

}  // namespace

extern "C" const char* pcr_geotiff_last_error(void) { return pcrb::g_geotiff_error.c_str(); }

namespace {

// Tiles of one image, band after band, appended to `f` at `pos`.  Gathering (and compression) of the tiles
// is spread over worker threads that claim tile indices from an atomic cursor and fill a ring of slots;
// this thread writes the slots to the file in tile order.  (zlib level 6 on float data runs at ~17 MB/s per
// core: one 1000 x 1000 band took 236 ms single-threaded.)
bool write_tiles(FILE* f, uint64_t& pos, const float* const* bands, int num_bands, int W, int H, int tw, int th,
                 int compression, int level, Image& img)
{
    img.W = W; img.H = H;
    const int tiles_x = (W + tw - 1) / tw, tiles_y = (H + th - 1) / th;
    const size_t per_band = static_cast<size_t>(tiles_x) * tiles_y;
    const size_t raw_tile = static_cast<size_t>(tw) * th * sizeof(float);
    const size_t total_tiles = per_band * static_cast<size_t>(num_bands);
    const float nan = std::numeric_limits<float>::quiet_NaN();
    const size_t zcap = compression == 8 ? compressBound(raw_tile) : 0;
    struct Slot { std::vector<float> tile; std::vector<uint8_t> z; size_t len = 0; std::atomic<uint64_t> ready{0}; };
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    // uncompressed small rasters are a few memcpys: not worth starting threads for
    const bool threaded = compression != 1 || static_cast<double>(total_tiles) * raw_tile >= 64e6;
    const size_t n_workers = threaded ? std::min<size_t>({size_t(hw), size_t(16), total_tiles}) : 1;
    const size_t n_slots = std::max<size_t>(2, n_workers * 2);
    std::vector<Slot> slots(n_slots);
    for (Slot& sl : slots) { sl.tile.resize(static_cast<size_t>(tw) * th); sl.z.resize(zcap); }
    std::atomic<size_t> cursor{0}, written{0};
    std::atomic<bool> failed{false};
    auto produce = [&](size_t t) {                               // tile t -> slot t % n_slots
        Slot& sl = slots[t % n_slots];
        while (t >= written.load(std::memory_order_acquire) + n_slots && !failed.load()) std::this_thread::yield();
        const int b = static_cast<int>(t / per_band);
        const size_t r = t % per_band;
        const int ty = static_cast<int>(r / tiles_x), tx = static_cast<int>(r % tiles_x);
        const float* src = bands[b];
        const int x0 = tx * tw, y0 = ty * th;
        const int cw = std::min(tw, W - x0), chh = std::min(th, H - y0);
        if (cw < tw || chh < th) std::fill(sl.tile.begin(), sl.tile.end(), nan);
        for (int row = 0; row < chh; ++row)
            std::memcpy(&sl.tile[static_cast<size_t>(row) * tw], src + static_cast<size_t>(y0 + row) * W + x0,
                        static_cast<size_t>(cw) * sizeof(float));
        sl.len = raw_tile;
        if (compression == 8) {
            uLongf zl = sl.z.size();
            if (compress2(sl.z.data(), &zl, reinterpret_cast<const Bytef*>(sl.tile.data()), raw_tile, level) != Z_OK)
                failed.store(true);
            sl.len = zl;
        } else if (compression == 5) {
            lzw_encode(reinterpret_cast<const uint8_t*>(sl.tile.data()), raw_tile, sl.z);
            sl.len = sl.z.size();
        }
        sl.ready.store(t + 1, std::memory_order_release);
    };
    std::vector<std::thread> workers;
    for (size_t w = 0; w + 1 < n_workers; ++w)
        workers.emplace_back([&] {
            for (size_t t; (t = cursor.fetch_add(1)) < total_tiles && !failed.load();) produce(t);
        });
    bool io_ok = true;
    for (size_t t = 0; t < total_tiles; ++t) {
        Slot& sl = slots[t % n_slots];
        if (workers.empty()) produce(cursor.fetch_add(1));       // single-core box: do it here
        while (sl.ready.load(std::memory_order_acquire) != t + 1 && !failed.load()) std::this_thread::yield();
        if (failed.load()) { io_ok = false; break; }
        const void* out = compression != 1 ? static_cast<const void*>(sl.z.data()) : static_cast<const void*>(sl.tile.data());
        img.tile_off.push_back(pos); img.tile_len.push_back(sl.len);
        if (io_ok) io_ok = std::fwrite(out, 1, sl.len, f) == sl.len;
        pos += sl.len;
        if (pos & 1) { std::fputc(0, f); ++pos; }                  // word alignment
        written.store(t + 1, std::memory_order_release);
        if (!io_ok) { failed.store(true); break; }
    }
    if (!io_ok) failed.store(true);
    written.store(total_tiles + n_slots, std::memory_order_release);   // release any waiting producer
    for (auto& w : workers) w.join();
    return io_ok;
}

int write_file(const char* path, const float* const* bands, int32_t num_bands, const pcr_grid_desc* grid,
               const char* const* band_names, int32_t epsg, const char* compress, int32_t compress_level,
               int32_t tile_width, int32_t tile_height, int32_t bigtiff, int32_t cloud_optimized)
{
    if (!path || !bands || !grid || num_bands <= 0)
        return fail(PCR_INVALID_ARGUMENT, "write_geotiff: bad arguments");
    const int W = grid->width, H = grid->height;
    if (W <= 0 || H <= 0) return fail(PCR_INVALID_ARGUMENT, "grid dimensions mismatch config");
    const std::string comp = compress ? compress : "NONE";
    int compression = 1;
    if (comp == "DEFLATE") compression = 8;
    else if (comp == "LZW") compression = 5;
    else if (!(comp == "NONE" || comp.empty()))
        return fail(PCR_NOT_IMPLEMENTED, "write_geotiff: compression '" + comp + "' is not supported (NONE, LZW, DEFLATE)");
    const int tw = tile_width > 0 ? (tile_width + 15) / 16 * 16 : 256;    // TIFF: tile dims multiple of 16
    const int th = tile_height > 0 ? (tile_height + 15) / 16 * 16 : 256;
    const size_t raw_tile = static_cast<size_t>(tw) * th * sizeof(float);
    const bool big = bigtiff != 0;
    {
        const double tiles = static_cast<double>((W + tw - 1) / tw) * ((H + th - 1) / th);
        if (!big && tiles * num_bands * raw_tile * (cloud_optimized ? 1.34 : 1.0) > 3.9e9)
            return fail(PCR_INVALID_ARGUMENT, "write_geotiff: raster exceeds classic TIFF 4 GB limit; set bigtiff");
    }
    const int level = std::max(1, std::min(9, compress_level));

    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(PCR_IO_ERROR, std::string("failed to create GeoTIFF: ") + path);

    // header (IFD offset patched at the end: the IFDs go after the pixel data)
    std::vector<uint8_t> hdr;
    hdr.push_back('I'); hdr.push_back('I');
    if (big) { put<uint16_t>(hdr, 43); put<uint16_t>(hdr, 8); put<uint16_t>(hdr, 0); put<uint64_t>(hdr, 0); }
    else     { put<uint16_t>(hdr, 42); put<uint32_t>(hdr, 0); }
    std::fwrite(hdr.data(), 1, hdr.size(), f);
    uint64_t pos = hdr.size();

    std::vector<Image> images(1);
    if (!write_tiles(f, pos, bands, num_bands, W, H, tw, th, compression, level, images[0])) {
        std::fclose(f);
        return fail(PCR_IO_ERROR, "failed to write band data");
    }
    // overview pyramid (write_geotiff with cloud_optimized, src/io/grid_io.cpp:155-176)
    if (cloud_optimized) {
        std::vector<std::vector<float>> cur(num_bands);
        int cw = W, chh = H;
        const int min_dim = std::min(W, H);
        for (int lvl = 2; min_dim / lvl >= 256; lvl *= 2) {
            std::vector<std::vector<float>> nxt(num_bands);
            int w2 = 0, h2 = 0;
            for (int b = 0; b < num_bands; ++b) {
                if (lvl == 2) {
                    std::vector<float> full(bands[b], bands[b] + static_cast<size_t>(W) * H);
                    halve(full, W, H, nxt[b], w2, h2);
                } else halve(cur[b], cw, chh, nxt[b], w2, h2);
            }
            cur.swap(nxt); cw = w2; chh = h2;
            std::vector<const float*> ptrs;
            for (auto& v : cur) ptrs.push_back(v.data());
            images.emplace_back();
            if (!write_tiles(f, pos, ptrs.data(), num_bands, cw, chh, tw, th, compression, level, images.back())) {
                std::fclose(f);
                return fail(PCR_IO_ERROR, "failed to write overview data");
            }
        }
    }

    const uint16_t nb = static_cast<uint16_t>(num_bands);
    const size_t inline_cap = big ? 8 : 4;
    std::vector<uint64_t> ifd_at(images.size(), 0);
    std::vector<std::vector<uint8_t>> ifds(images.size());
    for (size_t li = 0; li < images.size(); ++li) {
        const Image& im = images[li];
        std::vector<Entry> tags;
        if (li > 0) tags.push_back(longs(254, {1}));                       // NewSubfileType: reduced-resolution image
        tags.push_back(longs(256, {static_cast<uint32_t>(im.W)}));
        tags.push_back(longs(257, {static_cast<uint32_t>(im.H)}));
        tags.push_back(shorts(258, std::vector<uint16_t>(nb, 32)));
        tags.push_back(shorts(259, {static_cast<uint16_t>(compression)}));
        tags.push_back(shorts(262, {1}));                                  // MinIsBlack
        tags.push_back(shorts(277, {nb}));
        tags.push_back(shorts(284, {static_cast<uint16_t>(nb > 1 ? 2 : 1)}));  // planar: separate planes
        tags.push_back(longs(322, {static_cast<uint32_t>(tw)}));
        tags.push_back(longs(323, {static_cast<uint32_t>(th)}));
        tags.push_back(offsets(324, im.tile_off, big));
        tags.push_back(offsets(325, im.tile_len, big));
        if (nb > 1) tags.push_back(shorts(338, std::vector<uint16_t>(nb - 1, 0)));   // extra samples: unspecified
        tags.push_back(shorts(339, std::vector<uint16_t>(nb, 3)));                   // IEEE float
        tags.push_back(ascii(42113, "nan"));
        if (li == 0) {
            // georeferencing: GridConfig::gdal_geotransform = [min_x, csx, 0, max_y, 0, csy]
            const double ox = grid->min_x, oy = grid->max_y, sx = grid->cell_size_x, sy = grid->cell_size_y;
            if (sx > 0 && sy < 0) {
                tags.push_back(doubles(33550, {sx, -sy, 0.0}));
                tags.push_back(doubles(33922, {0, 0, 0, ox, oy, 0}));
            } else {
                tags.push_back(doubles(34264, {sx, 0, 0, ox, 0, sy, 0, oy, 0, 0, 0, 0, 0, 0, 0, 1}));
            }
            std::vector<uint16_t> keys = {1, 1, 0, 0};
            auto key = [&](uint16_t id, uint16_t value) { keys.insert(keys.end(), {id, 0, 1, value}); ++keys[3]; };
            const bool geographic = epsg >= 4000 && epsg < 5000;
            key(1024, epsg > 0 ? (geographic ? 2 : 1) : 32767);               // GTModelType
            key(1025, 1);                                                      // RasterPixelIsArea
            if (epsg > 0 && epsg < 65536) key(geographic ? 2048 : 3072, static_cast<uint16_t>(epsg));
            tags.push_back(shorts(34735, keys));
            std::string meta = "<GDALMetadata>\n";
            for (int b = 0; b < num_bands; ++b)
                if (band_names && band_names[b] && band_names[b][0])
                    meta += "  <Item name=\"DESCRIPTION\" sample=\"" + std::to_string(b) + "\" role=\"description\">" +
                            xml_escape(band_names[b]) + "</Item>\n";
            meta += "</GDALMetadata>\n";
            tags.push_back(ascii(42112, meta));
        }
        std::sort(tags.begin(), tags.end(), [](const Entry& a, const Entry& b) { return a.tag < b.tag; });

        // out-of-line payloads, then the IFD
        std::vector<uint64_t> where(tags.size(), 0);
        for (size_t i = 0; i < tags.size(); ++i) {
            if (tags[i].data.size() <= inline_cap) continue;
            where[i] = pos;
            std::fwrite(tags[i].data.data(), 1, tags[i].data.size(), f);
            pos += tags[i].data.size();
            if (pos & 1) { std::fputc(0, f); ++pos; }
        }
        std::vector<uint8_t>& ifd = ifds[li];
        if (big) put<uint64_t>(ifd, tags.size()); else put<uint16_t>(ifd, static_cast<uint16_t>(tags.size()));
        for (size_t i = 0; i < tags.size(); ++i) {
            put<uint16_t>(ifd, tags[i].tag);
            put<uint16_t>(ifd, tags[i].type);
            if (big) put<uint64_t>(ifd, tags[i].count); else put<uint32_t>(ifd, static_cast<uint32_t>(tags[i].count));
            std::vector<uint8_t> cell(inline_cap, 0);
            if (tags[i].data.size() <= inline_cap) std::memcpy(cell.data(), tags[i].data.data(), tags[i].data.size());
            else if (big) std::memcpy(cell.data(), &where[i], 8);
            else { const uint32_t o = static_cast<uint32_t>(where[i]); std::memcpy(cell.data(), &o, 4); }
            ifd.insert(ifd.end(), cell.begin(), cell.end());
        }
    }
    // IFD chain: main image first, then the overviews
    for (size_t li = 0; li < images.size(); ++li) {
        ifd_at[li] = pos;
        pos += ifds[li].size() + (big ? 8 : 4);
        if (pos & 1) ++pos;
    }
    if (!big && pos > 0xfffffff0ull) { std::fclose(f); return fail(PCR_INVALID_ARGUMENT, "classic TIFF overflow; set bigtiff"); }
    for (size_t li = 0; li < images.size(); ++li) {
        std::vector<uint8_t> blk = ifds[li];
        const uint64_t next = li + 1 < images.size() ? ifd_at[li + 1] : 0;
        if (big) put<uint64_t>(blk, next); else put<uint32_t>(blk, static_cast<uint32_t>(next));
        if (blk.size() & 1) blk.push_back(0);
        std::fwrite(blk.data(), 1, blk.size(), f);
    }
    std::fseek(f, big ? 8 : 4, SEEK_SET);
    if (big) std::fwrite(&ifd_at[0], 8, 1, f);
    else { const uint32_t o = static_cast<uint32_t>(ifd_at[0]); std::fwrite(&o, 4, 1, f); }
    if (std::fclose(f) != 0) return fail(PCR_IO_ERROR, "failed to close GeoTIFF");
    return PCR_OK;
}

template <typename F>
int guarded_io(F&& fn) noexcept
{
    try { return fn(); }
    catch (const std::bad_alloc&) { return fail(PCR_OUT_OF_MEMORY, "out of host memory"); }
    catch (const std::exception& e) { return fail(PCR_IO_ERROR, std::string("internal error: ") + e.what()); }
    catch (...) { return fail(PCR_IO_ERROR, "internal error"); }
}

// TiledGeoTiffWriter (include/pcr/io/grid_io.h:44-70): tiles arrive one at a time (out-of-core assembly);
// the raster is assembled in host memory and written, with its overviews, at close().
struct TiledWriter {
    std::string path, compress;
    pcr_grid_desc grid{};
    std::vector<std::string> names;
    std::vector<std::vector<float>> bands;
    int epsg = 0, level = 6, tw = 256, th = 256, bigtiff = 1, cog = 0;
};

}  // namespace

extern "C" int pcr_geotiff_write(const char* path, const float* const* bands, int32_t num_bands,
                                 const pcr_grid_desc* grid, const char* const* band_names, int32_t epsg,
                                 const char* compress, int32_t compress_level, int32_t tile_width,
                                 int32_t tile_height, int32_t bigtiff, int32_t cloud_optimized)
{
    return guarded_io([&] {
        return write_file(path, bands, num_bands, grid, band_names, epsg, compress, compress_level, tile_width,
                          tile_height, bigtiff, cloud_optimized);
    });
}

extern "C" int pcr_geotiff_tiled_open(const char* path, const pcr_grid_desc* grid, const char* const* band_names,
                                      int32_t num_bands, int32_t epsg, const char* compress, int32_t compress_level,
                                      int32_t tile_width, int32_t tile_height, int32_t bigtiff, int32_t cloud_optimized,
                                      void** handle)
{
    if (!handle) return fail(PCR_INVALID_ARGUMENT, "null handle pointer");
    *handle = nullptr;
    if (!path || !grid || num_bands <= 0 || grid->width <= 0 || grid->height <= 0 || grid->tile_width <= 0 || grid->tile_height <= 0)
        return fail(PCR_INVALID_ARGUMENT, "TiledGeoTiffWriter::open: bad arguments");
    return guarded_io([&] {
        TiledWriter* w = new TiledWriter();
        w->path = path; w->grid = *grid; w->epsg = epsg; w->compress = compress ? compress : "NONE";
        w->level = compress_level; w->tw = tile_width; w->th = tile_height; w->bigtiff = bigtiff; w->cog = cloud_optimized;
        for (int b = 0; b < num_bands; ++b) w->names.emplace_back(band_names && band_names[b] ? band_names[b] : "");
        w->bands.assign(num_bands, std::vector<float>(static_cast<size_t>(grid->width) * grid->height,
                                                     std::numeric_limits<float>::quiet_NaN()));
        *handle = w;
        return static_cast<int>(PCR_OK);
    });
}

// write_tile (src/io/grid_io.cpp:304-348): `data` is band-sequential, tile_cols x tile_rows floats per band,
// for reference tile (tile_row, tile_col) of the GridConfig (GridConfig::tile_cell_range)
extern "C" int pcr_geotiff_tiled_write_tile(void* handle, int32_t tile_row, int32_t tile_col, const float* data, int32_t num_bands)
{
    TiledWriter* w = static_cast<TiledWriter*>(handle);
    if (!w) return fail(PCR_INVALID_ARGUMENT, "writer not open");
    if (num_bands != static_cast<int32_t>(w->bands.size())) return fail(PCR_INVALID_ARGUMENT, "band count mismatch");
    if (!data) return fail(PCR_INVALID_ARGUMENT, "null data pointer");
    const pcr_grid_desc& g = w->grid;
    const int c0 = tile_col * g.tile_width, r0 = tile_row * g.tile_height;
    if (tile_row < 0 || tile_col < 0 || c0 >= g.width || r0 >= g.height) return fail(PCR_INVALID_ARGUMENT, "tile index out of range");
    const int cols = std::min(g.tile_width, g.width - c0), rows = std::min(g.tile_height, g.height - r0);
    for (int b = 0; b < num_bands; ++b)
        for (int r = 0; r < rows; ++r)
            std::memcpy(&w->bands[b][static_cast<size_t>(r0 + r) * g.width + c0],
                        data + (static_cast<size_t>(b) * rows + r) * cols, static_cast<size_t>(cols) * sizeof(float));
    return PCR_OK;
}

extern "C" int pcr_geotiff_tiled_close(void* handle)
{
    TiledWriter* w = static_cast<TiledWriter*>(handle);
    if (!w) return fail(PCR_INVALID_ARGUMENT, "writer not open");
    const int rc = guarded_io([&] {
        std::vector<const float*> ptrs;
        std::vector<const char*> names;
        for (auto& b : w->bands) ptrs.push_back(b.data());
        for (auto& n : w->names) names.push_back(n.c_str());
        return write_file(w->path.c_str(), ptrs.data(), static_cast<int32_t>(ptrs.size()), &w->grid, names.data(), w->epsg,
                          w->compress.c_str(), w->level, w->tw, w->th, w->bigtiff, w->cog);
    });
    delete w;
    return rc;
}

// Header reader: width, height, bands, EPSG and bounds of a (Big)TIFF written by
// the function above or by GDAL with the same georeferencing tags.
static int read_info_impl(const char* path, int32_t* width, int32_t* height, int32_t* num_bands,
                          int32_t* epsg, double bounds[4])
{
    if (!path || !width || !height || !num_bands || !epsg || !bounds)
        return fail(PCR_INVALID_ARGUMENT, "read_geotiff_info: bad arguments");
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail(PCR_IO_ERROR, std::string("failed to open file: ") + path);
    auto rd = [&](uint64_t off, void* dst, size_t n) {
        return std::fseek(f, static_cast<long>(off), SEEK_SET) == 0 && std::fread(dst, 1, n, f) == n;
    };
    uint8_t h[16];
    if (!rd(0, h, 8) || h[0] != 'I' || h[1] != 'I') { std::fclose(f); return fail(PCR_IO_ERROR, "not a little-endian TIFF"); }
    uint16_t magic; std::memcpy(&magic, h + 2, 2);
    const bool big = magic == 43;
    if (!big && magic != 42) { std::fclose(f); return fail(PCR_IO_ERROR, "not a TIFF file"); }
    uint64_t ifd = 0;
    if (big) { if (!rd(8, &ifd, 8)) { std::fclose(f); return fail(PCR_IO_ERROR, "truncated TIFF"); } }
    else { uint32_t o; std::memcpy(&o, h + 4, 4); ifd = o; }
    uint64_t n = 0;
    if (big) rd(ifd, &n, 8); else { uint16_t s = 0; rd(ifd, &s, 2); n = s; }
    std::fseek(f, 0, SEEK_END);
    const uint64_t file_size = static_cast<uint64_t>(std::ftell(f));
    if (n > file_size / 12) { std::fclose(f); return fail(PCR_IO_ERROR, "corrupt TIFF directory"); }
    const size_t esz = big ? 20 : 12, inl = big ? 8 : 4;
    *width = *height = 0; *num_bands = 1; *epsg = 0;
    std::vector<double> scale, tie, xform;
    std::vector<uint16_t> keys;
    for (uint64_t i = 0; i < n; ++i) {
        uint8_t e[20];
        if (!rd(ifd + (big ? 8 : 2) + i * esz, e, esz)) break;
        uint16_t tag, type; std::memcpy(&tag, e, 2); std::memcpy(&type, e + 2, 2);
        uint64_t count = 0;
        if (big) std::memcpy(&count, e + 4, 8); else { uint32_t c; std::memcpy(&c, e + 4, 4); count = c; }
        const size_t tsz = type == T_SHORT ? 2 : type == T_LONG ? 4 : (type == T_DOUBLE || type == T_LONG8) ? 8 : 1;
        if (count > file_size / tsz) continue;                     // an untrusted count must not size an allocation
        std::vector<uint8_t> payload(count * tsz);
        const uint8_t* cell = e + (big ? 12 : 8);
        if (payload.size() <= inl) std::memcpy(payload.data(), cell, payload.size());
        else {
            uint64_t off = 0;
            if (big) std::memcpy(&off, cell, 8); else { uint32_t o; std::memcpy(&o, cell, 4); off = o; }
            if (!rd(off, payload.data(), payload.size())) continue;
        }
        auto as_u = [&](size_t k) -> uint64_t {
            if (type == T_SHORT) { uint16_t v; std::memcpy(&v, &payload[k * 2], 2); return v; }
            if (type == T_LONG)  { uint32_t v; std::memcpy(&v, &payload[k * 4], 4); return v; }
            if (type == T_LONG8) { uint64_t v; std::memcpy(&v, &payload[k * 8], 8); return v; }
            return payload[k];
        };
        auto as_d = [&](std::vector<double>& out) {
            out.resize(count);
            if (type == T_DOUBLE) std::memcpy(out.data(), payload.data(), count * 8);
        };
        switch (tag) {
        case 256: *width = static_cast<int32_t>(as_u(0)); break;
        case 257: *height = static_cast<int32_t>(as_u(0)); break;
        case 277: *num_bands = static_cast<int32_t>(as_u(0)); break;
        case 33550: as_d(scale); break;
        case 33922: as_d(tie); break;
        case 34264: as_d(xform); break;
        case 34735: keys.resize(count); for (size_t k = 0; k < count; ++k) keys[k] = static_cast<uint16_t>(as_u(k)); break;
        default: break;
        }
    }
    std::fclose(f);
    for (size_t k = 4; k + 3 < keys.size(); k += 4)
        if ((keys[k] == 3072 || keys[k] == 2048) && keys[k + 1] == 0) *epsg = keys[k + 3];
    double gt[6] = {0, 1, 0, 0, 0, 1};
    if (scale.size() >= 2 && tie.size() >= 6) {
        gt[0] = tie[3] - tie[0] * scale[0]; gt[1] = scale[0]; gt[3] = tie[4] + tie[1] * scale[1]; gt[5] = -scale[1];
    } else if (xform.size() >= 8) {
        gt[0] = xform[3]; gt[1] = xform[0]; gt[3] = xform[7]; gt[5] = xform[5];
    }
    // same arithmetic as read_geotiff_info, src/io/grid_io.cpp:411-421
    bounds[0] = gt[0];                       // min_x
    bounds[3] = gt[3];                       // max_y
    bounds[2] = gt[0] + gt[1] * *width;      // max_x
    bounds[1] = gt[3] + gt[5] * *height;     // min_y
    return PCR_OK;
}

extern "C" int pcr_geotiff_read_info(const char* path, int32_t* width, int32_t* height, int32_t* num_bands,
                                     int32_t* epsg, double bounds[4])
{
    return guarded_io([&] { return read_info_impl(path, width, height, num_bands, epsg, bounds); });
}

// read_geotiff_band (src/io/grid_io.cpp:445-497) for files of this writer's layout: tiled, Float32, one plane
// per band, compression NONE / LZW / DEFLATE; the first (full-resolution) image.
extern "C" int pcr_geotiff_read_band(const char* path, int32_t band_index, float* data, int32_t width, int32_t height)
{
    if (!data) return fail(PCR_INVALID_ARGUMENT, "null data pointer");
    if (band_index < 0) return fail(PCR_INVALID_ARGUMENT, "invalid band index");
    return guarded_io([&]() -> int {
        FILE* f = std::fopen(path, "rb");
        if (!f) return fail(PCR_IO_ERROR, std::string("failed to open file: ") + path);
        struct Closer { FILE* f; ~Closer() { std::fclose(f); } } closer{f};
        std::fseek(f, 0, SEEK_END);
        const uint64_t file_size = static_cast<uint64_t>(std::ftell(f));
        auto rd = [&](uint64_t off, void* dst, size_t n) {
            return off + n <= file_size && std::fseek(f, static_cast<long>(off), SEEK_SET) == 0 && std::fread(dst, 1, n, f) == n;
        };
        uint8_t h[16];
        if (!rd(0, h, 8) || h[0] != 'I' || h[1] != 'I') return fail(PCR_IO_ERROR, "not a little-endian TIFF");
        uint16_t magic; std::memcpy(&magic, h + 2, 2);
        const bool big = magic == 43;
        if (!big && magic != 42) return fail(PCR_IO_ERROR, "not a TIFF file");
        uint64_t ifd = 0;
        if (big) { if (!rd(8, &ifd, 8)) return fail(PCR_IO_ERROR, "truncated TIFF"); }
        else { uint32_t o; std::memcpy(&o, h + 4, 4); ifd = o; }
        uint64_t n = 0;
        if (big) rd(ifd, &n, 8); else { uint16_t s = 0; rd(ifd, &s, 2); n = s; }
        if (n > file_size / 12) return fail(PCR_IO_ERROR, "corrupt TIFF directory");
        const size_t esz = big ? 20 : 12, inl = big ? 8 : 4;
        uint64_t W = 0, H = 0, nb = 1, tw = 0, th = 0, comp = 1, planar = 1;
        std::vector<uint64_t> offs, lens;
        for (uint64_t i = 0; i < n; ++i) {
            uint8_t e[20];
            if (!rd(ifd + (big ? 8 : 2) + i * esz, e, esz)) break;
            uint16_t tag, type; std::memcpy(&tag, e, 2); std::memcpy(&type, e + 2, 2);
            uint64_t count = 0;
            if (big) std::memcpy(&count, e + 4, 8); else { uint32_t c; std::memcpy(&c, e + 4, 4); count = c; }
            const size_t tsz = type == T_SHORT ? 2 : type == T_LONG ? 4 : (type == T_DOUBLE || type == T_LONG8) ? 8 : 1;
            if (count > file_size / tsz) continue;
            std::vector<uint8_t> payload(count * tsz);
            const uint8_t* cell = e + (big ? 12 : 8);
            if (payload.size() <= inl) std::memcpy(payload.data(), cell, payload.size());
            else {
                uint64_t off = 0;
                if (big) std::memcpy(&off, cell, 8); else { uint32_t o; std::memcpy(&o, cell, 4); off = o; }
                if (!rd(off, payload.data(), payload.size())) continue;
            }
            auto as_u = [&](size_t k) -> uint64_t {
                if (type == T_SHORT) { uint16_t v; std::memcpy(&v, &payload[k * 2], 2); return v; }
                if (type == T_LONG)  { uint32_t v; std::memcpy(&v, &payload[k * 4], 4); return v; }
                if (type == T_LONG8) { uint64_t v; std::memcpy(&v, &payload[k * 8], 8); return v; }
                return payload[k];
            };
            auto all = [&](std::vector<uint64_t>& out) { out.resize(count); for (size_t k = 0; k < count; ++k) out[k] = as_u(k); };
            switch (tag) {
            case 256: W = as_u(0); break;
            case 257: H = as_u(0); break;
            case 259: comp = as_u(0); break;
            case 277: nb = as_u(0); break;
            case 284: planar = as_u(0); break;
            case 322: tw = as_u(0); break;
            case 323: th = as_u(0); break;
            case 324: all(offs); break;
            case 325: all(lens); break;
            default: break;
            }
        }
        if (static_cast<int64_t>(W) != width || static_cast<int64_t>(H) != height) return fail(PCR_INVALID_ARGUMENT, "dimension mismatch");
        if (static_cast<uint64_t>(band_index) >= nb) return fail(PCR_INVALID_ARGUMENT, "band index out of range");
        if (tw == 0 || th == 0 || (nb > 1 && planar != 2) || !(comp == 1 || comp == 5 || comp == 8))
            return fail(PCR_NOT_IMPLEMENTED, "read_geotiff_band: only tiled, band-separate NONE/LZW/DEFLATE files are supported");
        const uint64_t tiles_x = (W + tw - 1) / tw, tiles_y = (H + th - 1) / th, per_band = tiles_x * tiles_y;
        if (offs.size() < per_band * nb || lens.size() < per_band * nb) return fail(PCR_IO_ERROR, "corrupt tile tables");
        const size_t raw = tw * th * sizeof(float);
        std::vector<uint8_t> zbuf;
        std::vector<float> tile(tw * th);
        for (uint64_t t = 0; t < per_band; ++t) {
            const uint64_t k = static_cast<uint64_t>(band_index) * per_band + t;
            if (lens[k] > file_size) return fail(PCR_IO_ERROR, "corrupt tile length");
            zbuf.resize(lens[k]);
            if (!rd(offs[k], zbuf.data(), zbuf.size())) return fail(PCR_IO_ERROR, "failed to read band data");
            if (comp == 1) { if (zbuf.size() != raw) return fail(PCR_IO_ERROR, "corrupt tile"); std::memcpy(tile.data(), zbuf.data(), raw); }
            else if (comp == 8) {
                uLongf out = raw;
                if (uncompress(reinterpret_cast<Bytef*>(tile.data()), &out, zbuf.data(), zbuf.size()) != Z_OK || out != raw)
                    return fail(PCR_IO_ERROR, "failed to inflate tile");
            } else if (!lzw_decode(zbuf.data(), zbuf.size(), reinterpret_cast<uint8_t*>(tile.data()), raw))
                return fail(PCR_IO_ERROR, "failed to decode LZW tile");
            const uint64_t ty = t / tiles_x, tx = t % tiles_x, y0 = ty * th, x0 = tx * tw;
            const uint64_t cw = std::min<uint64_t>(tw, W - x0), chh = std::min<uint64_t>(th, H - y0);
            for (uint64_t r = 0; r < chh; ++r)
                std::memcpy(data + (y0 + r) * W + x0, &tile[r * tw], cw * sizeof(float));
        }
        return static_cast<int>(PCR_OK);
    });
}
