// checkpoint.cu — tile-state checkpoints in the reference's .pcrt format (host-side C++).
//
// Reference: write_tile_state / read_tile_state / tile_state_filename
// (src/io/tile_state_io.cpp:14-211) and the reload rule of TileManager::acquire
// (src/engine/tile_manager.cpp:272-302).  The reference's state is band-sequential float per
// reduction per tile; ours is one interleaved record per cell shared by the fused reductions, so
// save/load convert between the two on the host: records are copied D2H once, each reduction's
// state floats are gathered per tile (max/min words decoded from the ordered-int map) and written;
// load does the reverse and uploads.
#include "engine.h"

#include <sys/stat.h>

#include <cstdio>
#include <cstring>

namespace pcrb {

#define CU_TRY(expr)                                                                     \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess)                                                           \
            return Status::error(PCR_CUDA_ERROR, std::string("CUDA error: ") +           \
                                 cudaGetErrorString(_e) + " (" #expr ")");               \
    } while (0)
#define ST_TRY(expr) do { Status _s = (expr); if (!_s.ok()) return _s; } while (0)

namespace {

#pragma pack(push, 1)
struct PcrtHeader {            // 36 bytes, little-endian
    uint32_t magic;            // "PCRT"
    uint32_t version;          // 1
    int32_t tile_row, tile_col, cols, rows, state_floats;
    uint8_t reduction;
    uint8_t reserved[7];
};
#pragma pack(pop)
static_assert(sizeof(PcrtHeader) == 36, "pcrt header is 36 bytes");
constexpr uint32_t kMagic = 0x54524350u;

std::string tile_path(const std::string& dir, int row, int col)
{
    char name[64];
    std::snprintf(name, sizeof name, "tile_%04d_%04d.pcrt", row, col);
    return (dir.empty() || dir.back() == '/') ? dir + name : dir + "/" + name;
}

bool make_dir(const std::string& d)
{
    struct stat st;
    if (stat(d.c_str(), &st) == 0) return S_ISDIR(st.st_mode);
    return mkdir(d.c_str(), 0777) == 0;
}

}  // namespace

// Which record words hold reduction `band`'s state floats, in the reference's order
// (Sum: {sum}; Count: {count}; Average/WeightedAverage: {sum, count}; Max/Min: {value}).
static bool band_words(const Pass& p, int band, int& kind, int& wa, int& wb)
{
    for (int i = 0; i < p.fin.n; ++i)
        if (p.fin.band[i] == band) { kind = p.fin.kind[i]; wa = p.fin.word_a[i]; wb = p.fin.word_b[i]; return true; }
    return false;
}

// Records of one reference tile, D2H / H2D with a pitched copy: the host never holds more than one
// tile (<= tile_w * tile_h records; 268 MB for a 4096^2 tile of 16-byte records) instead of the whole grid.
static cudaError_t copy_tile(uint32_t* host, uint32_t* dev_state, const GridParams& g, int W, int c0, int r0,
                             int cols, int rows, bool to_host)
{
    uint32_t* d = dev_state + (static_cast<size_t>(r0) * g.width + c0) * W;
    const size_t dpitch = static_cast<size_t>(g.width) * W * 4, hpitch = static_cast<size_t>(cols) * W * 4;
    return to_host ? cudaMemcpy2D(host, hpitch, d, dpitch, hpitch, rows, cudaMemcpyDeviceToHost)
                   : cudaMemcpy2D(d, dpitch, host, hpitch, hpitch, rows, cudaMemcpyHostToDevice);
}

Status Engine::save_state(const std::string& dir)
{
    CU_TRY(cudaSetDevice(device_));
    // N>1: a rank's records are only its own contribution since the last finalize, and the merged state
    // lives in row slices spread over the ranks; the reference's per-tile files cannot express either.
    if (exact_)
        return Status::error(PCR_NOT_IMPLEMENTED, "pipeline: save_state is not available in deterministic mode 2 (exact fixed-point state)");
    if (world_ > 1)
        return Status::error(PCR_NOT_IMPLEMENTED,
                             "pipeline: save_state is not available on a multi-GPU pipeline (the accumulated state "
                             "is distributed over the ranks' row slices); checkpoint from a single-GPU pipeline");
    ST_TRY(bin_flush_all());
    ST_TRY(synchronize());
    if (!make_dir(dir)) return Status::error(PCR_IO_ERROR, "failed to create state directory: " + dir);
    std::vector<uint32_t> touched(std::max(1, n_tiles_));
    CU_TRY(cudaMemcpy(touched.data(), d_touched_, touched.size() * 4, cudaMemcpyDeviceToHost));
    const bool single = reductions_.size() == 1;
    std::vector<uint32_t> rec;
    std::vector<float> out;
    for (Pass& p : passes_) {
        const int W = p.layout.width;
        for (int i = 0; i < p.fin.n; ++i) {
            const std::string bdir = single ? dir : dir + "/band_" + std::to_string(p.fin.band[i]);
            if (!make_dir(bdir)) return Status::error(PCR_IO_ERROR, "failed to create state directory: " + bdir);
        }
        for (int ty = 0; ty < gp_.tiles_y; ++ty)
            for (int tx = 0; tx < gp_.tiles_x; ++tx) {
                if (!touched[ty * gp_.tiles_x + tx]) continue;
                const int c0 = tx * gp_.tile_w, r0 = ty * gp_.tile_h;
                const int cols = std::min(gp_.tile_w, gp_.width - c0), rows = std::min(gp_.tile_h, gp_.height - r0);
                const size_t tc = static_cast<size_t>(cols) * rows;
                rec.resize(tc * W);
                CU_TRY(copy_tile(rec.data(), p.d_state, gp_, W, c0, r0, cols, rows, true));
                for (int i = 0; i < p.fin.n; ++i) {
                    const int band = p.fin.band[i];
                    int kind, wa, wb;
                    band_words(p, band, kind, wa, wb);
                    const int sf = kind == FIN_RATIO ? 2 : 1;
                    const std::string bdir = single ? dir : dir + "/band_" + std::to_string(band);
                    out.resize(tc * sf);
                    auto as_f = [](uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; };
                    for (size_t li = 0; li < tc; ++li) {
                        const uint32_t* q = &rec[li * W];
                        if (kind == FIN_MAX || kind == FIN_MIN) out[li] = ordered_f32(static_cast<int32_t>(q[wa]));
                        else out[li] = as_f(q[wa]);
                        if (sf == 2) out[tc + li] = as_f(q[wb]);
                    }
                    PcrtHeader h{};
                    h.magic = kMagic; h.version = 1; h.tile_row = ty; h.tile_col = tx; h.cols = cols; h.rows = rows;
                    h.state_floats = sf; h.reduction = static_cast<uint8_t>(reductions_[band].type);
                    const std::string path = tile_path(bdir, ty, tx);
                    FILE* f = std::fopen(path.c_str(), "wb");
                    if (!f) return Status::error(PCR_IO_ERROR, "failed to open file for writing: " + path);
                    const bool ok = std::fwrite(&h, sizeof h, 1, f) == 1 && std::fwrite(out.data(), 4, out.size(), f) == out.size();
                    if (std::fclose(f) != 0 || !ok) return Status::error(PCR_IO_ERROR, "failed to write state data");
                }
            }
    }
    return Status::success();
}

Status Engine::load_state(const std::string& dir)
{
    CU_TRY(cudaSetDevice(device_));
    ST_TRY(bin_flush_all());
    ST_TRY(synchronize());
    // N>1: the files are ONE contribution to the merged state; rank 0 takes them into its partial state
    // (merged into the owners' slices at the next finalize), the other ranks keep the identity — loading
    // them everywhere would count the checkpoint world-size times.
    if (partition_ || exact_)
        return Status::error(PCR_NOT_IMPLEMENTED, "pipeline: load_state is not available with the tile-partitioned multi-GPU layout "
                                                  "or in deterministic mode 2");
    if (world_ > 1 && rank_ != 0) return Status::success();
    std::vector<uint32_t> touched(std::max(1, n_tiles_));
    CU_TRY(cudaMemcpy(touched.data(), d_touched_, touched.size() * 4, cudaMemcpyDeviceToHost));
    const bool single = reductions_.size() == 1;
    std::vector<uint32_t> rec;
    std::vector<float> in;
    for (Pass& p : passes_) {
        const int W = p.layout.width;
        for (int ty = 0; ty < gp_.tiles_y; ++ty)
            for (int tx = 0; tx < gp_.tiles_x; ++tx) {
                const int c0 = tx * gp_.tile_w, r0 = ty * gp_.tile_h;
                const int cols = std::min(gp_.tile_w, gp_.width - c0), rows = std::min(gp_.tile_h, gp_.height - r0);
                const size_t tc = static_cast<size_t>(cols) * rows;
                bool have_rec = false, changed = false;
                for (int i = 0; i < p.fin.n; ++i) {
                    const int band = p.fin.band[i];
                    int kind, wa, wb;
                    band_words(p, band, kind, wa, wb);
                    const int sf = kind == FIN_RATIO ? 2 : 1;
                    const std::string bdir = single ? dir : dir + "/band_" + std::to_string(band);
                    const std::string path = tile_path(bdir, ty, tx);
                    FILE* f = std::fopen(path.c_str(), "rb");
                    if (!f) continue;                                   // no file: the tile keeps its state
                    PcrtHeader h{};
                    const bool hok = std::fread(&h, sizeof h, 1, f) == 1 && h.magic == kMagic && h.version == 1;
                    // the reference discards files whose dimensions do not match (tile_manager.cpp:283-302)
                    if (!hok || h.cols != cols || h.rows != rows || h.state_floats != sf || h.tile_row != ty || h.tile_col != tx) {
                        std::fclose(f);
                        continue;
                    }
                    // a Max file must not land in a Min band (Average and WeightedAverage share one state layout
                    // and one meaning: {sum, count})
                    auto family = [](int t) { return t == PCR_WEIGHTED_AVERAGE ? static_cast<int>(PCR_AVERAGE) : t; };
                    if (family(h.reduction) != family(reductions_[band].type)) {
                        std::fclose(f);
                        return Status::error(PCR_INVALID_ARGUMENT, "state file holds a different reduction type than band " +
                                                                       std::to_string(band) + ": " + path);
                    }
                    in.resize(tc * sf);
                    const bool dok = std::fread(in.data(), 4, in.size(), f) == in.size();
                    std::fclose(f);
                    if (!dok) return Status::error(PCR_IO_ERROR, "incomplete state data (file truncated?): " + path);
                    if (!have_rec) {
                        rec.resize(tc * W);
                        CU_TRY(copy_tile(rec.data(), p.d_state, gp_, W, c0, r0, cols, rows, true));
                        have_rec = true;
                    }
                    auto as_u = [](float v) { uint32_t u; std::memcpy(&u, &v, 4); return u; };
                    for (size_t li = 0; li < tc; ++li) {
                        uint32_t* q = &rec[li * W];
                        if (kind == FIN_MAX || kind == FIN_MIN) q[wa] = static_cast<uint32_t>(f32_ordered(in[li]));
                        else q[wa] = as_u(in[li]);
                        if (sf == 2) q[wb] = as_u(in[tc + li]);
                    }
                    touched[ty * gp_.tiles_x + tx] = 1;
                    changed = true;
                }
                if (changed) CU_TRY(copy_tile(rec.data(), p.d_state, gp_, W, c0, r0, cols, rows, false));
            }
    }
    CU_TRY(cudaMemcpy(d_touched_, touched.data(), touched.size() * 4, cudaMemcpyHostToDevice));
    finalized_ = false;
    return Status::success();
}

}  // namespace pcrb
