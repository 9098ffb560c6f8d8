/*
 * pcr_oracle.c — TEST INFRASTRUCTURE.  CPU restatement (plain C11, scalar, one
 * thread) of the reference's ingest/finalize arithmetic.  It exists only so the
 * CUDA path can be checked against it: nothing under pointcloud_raster_b200/
 * may include, link or call this file.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here
 * against (a) the known-answer vectors of the reference's own gtests
 * (tests/cpp/test_grid_config.cpp, test_tile_router.cpp, test_accumulator.cpp,
 * test_reduction_ops.cpp, test_pipeline.cpp) and (b) outputs of the reference
 * itself, compiled unmodified into oracle/_ref (oracle/Makefile) and run in the
 * build container; the vectors (b) are committed under tests/golden/ together
 * with the script that made them (oracle/make_golden.py).
 *
 * All file:line citations are relative to /root/reference.
 *
 * State layout: one full-grid, band-sequential array per reduction
 * (state[f*cells + cell], include/pcr/ops/builtin_ops.h:143-176).  The
 * reference keeps the same floats split per 4096^2 tile; the split only
 * matters through two rules that are restated explicitly below:
 *   - the "touched tile" rule  (src/engine/tile_manager.cpp:437-444,
 *     src/engine/pipeline.cpp:1204-1226): untouched tiles finalize to NaN;
 *   - glyph clipping to the tile that holds the point's routed cell
 *     (src/engine/pipeline.cpp:699-709, glyph_kernels.cu:151-154,266-267).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off (no FMA contraction: the Line
 * endpoint and Gaussian weight expressions must round the way the reference's
 * ISO-mode x86-64 build does).
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* Values of pcr::ReductionType (include/pcr/core/types.h:34-46). */
enum { ORC_SUM = 0, ORC_MAX = 1, ORC_MIN = 2, ORC_AVERAGE = 3,
       ORC_WEIGHTED_AVERAGE = 4, ORC_COUNT = 5 };
/* Values of pcr::GlyphType (include/pcr/engine/glyph.h:10-14). */
enum { ORC_GLYPH_POINT = 0, ORC_GLYPH_LINE = 1, ORC_GLYPH_GAUSSIAN = 2 };

typedef struct {
    double min_x, min_y, max_x, max_y;   /* BBox */
    double cell_size_x, cell_size_y;     /* cell_size_y < 0 for north-up */
    int32_t width, height;               /* cells */
    int32_t tile_width, tile_height;     /* cells per tile (default 4096) */
} orc_grid;

typedef struct {
    int32_t type;        /* ORC_* reduction */
    int32_t glyph;       /* ORC_GLYPH_* */
    const float *value;  /* value channel, n floats */
    const float *direction, *half_length;          /* Line; NULL => default */
    const float *sigma_x, *sigma_y, *rotation;     /* Gaussian; NULL => default */
    float default_direction, default_half_length;
    float default_sigma_x, default_sigma_y, default_rotation;
    float max_radius_cells;
} orc_reduction;

/* ------------------------------------------------------------------------ */
/* Grid geometry                                                             */
/* ------------------------------------------------------------------------ */

/* GridConfig::compute_dimensions, src/core/grid_config.cpp:7-22. */
void orc_compute_dimensions(orc_grid *g)
{
    if (!(g->max_x >= g->min_x && g->max_y >= g->min_y)) {
        g->width = g->height = 0;
        return;
    }
    g->width  = (int32_t)ceil((g->max_x - g->min_x) / fabs(g->cell_size_x));
    g->height = (int32_t)ceil((g->max_y - g->min_y) / fabs(g->cell_size_y));
}

/* GridConfig::world_to_cell, src/core/grid_config.cpp:24-43, with
 * BBox::contains, src/core/types.cpp:41-43 (inclusive on all four edges; any
 * NaN coordinate fails every comparison and is rejected).  IEEE f64 division,
 * floor, then clamp into [0,w-1] x [0,h-1]. */
int orc_world_to_cell(const orc_grid *g, double wx, double wy,
                      int32_t *col, int32_t *row)
{
    if (!(wx >= g->min_x && wx <= g->max_x && wy >= g->min_y && wy <= g->max_y))
        return 0;
    int32_t c = (int32_t)floor((wx - g->min_x) / g->cell_size_x);
    int32_t r = (int32_t)floor((wy - g->max_y) / g->cell_size_y);
    if (c > g->width - 1)  c = g->width - 1;
    if (c < 0)             c = 0;
    if (r > g->height - 1) r = g->height - 1;
    if (r < 0)             r = 0;
    *col = c;
    *row = r;
    return 1;
}

static int32_t tiles_x_of(const orc_grid *g)
{
    return (g->width + g->tile_width - 1) / g->tile_width;
}
static int32_t tiles_y_of(const orc_grid *g)
{
    return (g->height + g->tile_height - 1) / g->tile_height;
}

/* TileRouter::assign (CPU), src/engine/tile_router.cpp:90-123: per point the
 * global cell (row*width+col, u32), the tile (row/th)*tiles_x + col/tw, and a
 * valid flag.  Invalid points get cell = tile = 0, valid = 0. */
void orc_assign(const orc_grid *g, const double *x, const double *y, size_t n,
                uint32_t *cell, uint32_t *tile, uint8_t *valid)
{
    const int32_t tx = tiles_x_of(g);
    for (size_t i = 0; i < n; ++i) {
        int32_t c, r;
        if (!orc_world_to_cell(g, x[i], y[i], &c, &r)) {
            cell[i] = 0; tile[i] = 0; valid[i] = 0;
            continue;
        }
        valid[i] = 1;
        cell[i]  = (uint32_t)(r * g->width + c);
        tile[i]  = (uint32_t)((r / g->tile_height) * tx + c / g->tile_width);
    }
}

/* ------------------------------------------------------------------------ */
/* Reducer algebra (include/pcr/ops/builtin_ops.h:10-103)                    */
/* ------------------------------------------------------------------------ */

int orc_state_floats(int type)
{
    return (type == ORC_AVERAGE || type == ORC_WEIGHTED_AVERAGE) ? 2 : 1;
}

/* init_state_cpu, src/ops/reduction_registry.cpp:28-40: identity per op. */
void orc_state_init(int type, float *state, int64_t cells)
{
    float id = 0.0f;
    if (type == ORC_MAX) id = -FLT_MAX;
    if (type == ORC_MIN) id =  FLT_MAX;
    const int64_t n = cells * orc_state_floats(type);
    for (int64_t i = 0; i < n; ++i) state[i] = id;
}

/* Op::combine for the Point glyph, builtin_ops.h:13,26,39,52,65,86-88.
 * WeightedAverage uses the unweighted combine (weight 1): the pipeline never
 * passes weights (src/engine/pipeline.cpp:669,679). */
static void combine_point(int type, float *state, int64_t cells, int64_t cell,
                          float v)
{
    switch (type) {
    case ORC_SUM:   state[cell] = state[cell] + v;         break;
    case ORC_MAX:   state[cell] = fmaxf(state[cell], v);   break;
    case ORC_MIN:   state[cell] = fminf(state[cell], v);   break;
    case ORC_COUNT: state[cell] = state[cell] + 1.0f;      break;
    case ORC_AVERAGE:
    case ORC_WEIGHTED_AVERAGE:
        state[cell]         = state[cell] + v;
        state[cells + cell] = state[cells + cell] + 1.0f;
        break;
    default: break;
    }
}

/* update_state_cpu, src/engine/glyph_kernels.cu:36-74: glyph contribution of
 * weight w.  Max/Min are rejected before this point (pipeline.cpp:500-508). */
static void combine_glyph(int type, float *state, int64_t cells, int64_t cell,
                          float v, float w)
{
    switch (type) {
    case ORC_AVERAGE:
    case ORC_WEIGHTED_AVERAGE:
        state[cell]         += v * w;
        state[cells + cell] += w;
        break;
    case ORC_SUM:   state[cell] += v * w; break;
    case ORC_COUNT: state[cell] += w;     break;
    default: break;
    }
}

/* Op::merge, builtin_ops.h:15,28,41,54,67,95-97 (the multi-GPU combine rule). */
void orc_state_merge(int type, float *dst, const float *src, int64_t cells)
{
    const int64_t n = cells * orc_state_floats(type);
    for (int64_t i = 0; i < n; ++i) {
        if (type == ORC_MAX)      dst[i] = fmaxf(dst[i], src[i]);
        else if (type == ORC_MIN) dst[i] = fminf(dst[i], src[i]);
        else                      dst[i] = dst[i] + src[i];
    }
}

/* Op::finalize, builtin_ops.h:16,29,42,55,68-70,99-101. */
static float finalize_cell(int type, const float *state, int64_t cells,
                           int64_t cell)
{
    const float a = state[cell];
    switch (type) {
    case ORC_SUM:   return a;
    case ORC_MAX:   return a == -FLT_MAX ? NAN : a;
    case ORC_MIN:   return a ==  FLT_MAX ? NAN : a;
    case ORC_COUNT: return a > 0.0f ? a : NAN;
    case ORC_AVERAGE:
    case ORC_WEIGHTED_AVERAGE: {
        const float b = state[cells + cell];
        return b > 0.0f ? a / b : NAN;
    }
    default: return NAN;
    }
}

/* Pipeline::Impl::finalize_result, src/engine/pipeline.cpp:1186-1281: band is
 * NaN everywhere, then every tile that has state is finalized cell by cell.
 * touched[tile] restates TileManager::tile_has_state (tile_manager.cpp:437-444)
 * for a run that starts from an empty state_dir. */
void orc_finalize(const orc_grid *g, int type, const float *state,
                  const uint8_t *touched, float *out)
{
    const int64_t cells = (int64_t)g->width * g->height;
    const int32_t tx = tiles_x_of(g);
    for (int32_t r = 0; r < g->height; ++r) {
        for (int32_t c = 0; c < g->width; ++c) {
            const int64_t cell = (int64_t)r * g->width + c;
            const int32_t t = (r / g->tile_height) * tx + c / g->tile_width;
            out[cell] = touched[t] ? finalize_cell(type, state, cells, cell) : NAN;
        }
    }
}

/* ------------------------------------------------------------------------ */
/* Glyph footprints                                                          */
/* ------------------------------------------------------------------------ */

typedef struct { int32_t c0, r0, w, h; } tile_rect;

/* GridConfig::tile_cell_range, src/core/grid_config.cpp:81-91, for the tile
 * that holds routed cell (col,row). */
static tile_rect tile_of(const orc_grid *g, int32_t col, int32_t row)
{
    tile_rect t;
    t.c0 = (col / g->tile_width)  * g->tile_width;
    t.r0 = (row / g->tile_height) * g->tile_height;
    t.w  = g->tile_width  < g->width  - t.c0 ? g->tile_width  : g->width  - t.c0;
    t.h  = g->tile_height < g->height - t.r0 ? g->tile_height : g->height - t.r0;
    return t;
}

/* A footprint visitor receives every (global cell, weight) pair a glyph paints,
 * in the reference's loop order. */
typedef void (*orc_visit)(void *ctx, int64_t cell, float w);

/* std::min / std::max on floats as libstdc++ defines them — the reference calls
 * std::min(a,b) = (b<a)?b:a and std::max(a,b) = (a<b)?b:a
 * (glyph_kernels.cu:131,228-229); kept distinct from fminf/fmaxf because the
 * NaN behaviour differs. */
static float std_minf(float a, float b) { return (b < a) ? b : a; }
static float std_maxf(float a, float b) { return (a < b) ? b : a; }

/* accumulate_glyph_line_cpu, src/engine/glyph_kernels.cu:188-281. */
static void footprint_line(const orc_grid *g, const orc_reduction *rd, size_t p,
                           double wx, double wy, tile_rect t,
                           orc_visit visit, void *ctx)
{
    const double inv_csx = 1.0 / g->cell_size_x;
    const double inv_csy = 1.0 / g->cell_size_y;
    const double fcx = (wx - g->min_x) * inv_csx;     /* multiply, not divide */
    const double fcy = (wy - g->max_y) * inv_csy;

    const float dir = rd->direction   ? rd->direction[p]   : rd->default_direction;
    const float hl  = rd->half_length ? rd->half_length[p] : rd->default_half_length;

    /* world -> cells; with cell_size_y < 0, hy is negative and the cap never
     * binds on it. */
    const float cap = rd->max_radius_cells;
    const float hx = std_minf(hl * (float)inv_csx, cap);
    const float hy = std_minf(hl * (float)inv_csy, cap);

    /* float product, then f64 subtract/add, then round half away from zero. */
    const float px = hx * cosf(dir);
    const float py = hy * sinf(dir);
    const int32_t ix0 = (int32_t)round(fcx - (double)px);
    const int32_t iy0 = (int32_t)round(fcy - (double)py);
    const int32_t ix1 = (int32_t)round(fcx + (double)px);
    const int32_t iy1 = (int32_t)round(fcy + (double)py);

    const int32_t adx = ix1 > ix0 ? ix1 - ix0 : ix0 - ix1;
    const int32_t ady = iy1 > iy0 ? iy1 - iy0 : iy0 - iy1;
    const int32_t stepx = ix0 < ix1 ? 1 : -1;
    const int32_t stepy = iy0 < iy1 ? 1 : -1;
    int32_t err = adx - ady;
    int32_t cx = ix0, cy = iy0;
    const int32_t max_steps = 2 * (adx + ady) + 2;

    for (int32_t s = 0; s <= max_steps; ++s) {
        const int32_t lc = cx - t.c0, lr = cy - t.r0;
        if (lc >= 0 && lc < t.w && lr >= 0 && lr < t.h)
            visit(ctx, (int64_t)cy * g->width + cx, 1.0f);
        if (cx == ix1 && cy == iy1) break;
        const int32_t e2 = 2 * err;
        if (e2 > -ady) { err -= ady; cx += stepx; }
        if (e2 <  adx) { err += adx; cy += stepy; }
    }
}

/* accumulate_glyph_gaussian_cpu, src/engine/glyph_kernels.cu:79-183. */
static void footprint_gaussian(const orc_grid *g, const orc_reduction *rd,
                               size_t p, double wx, double wy, tile_rect t,
                               orc_visit visit, void *ctx)
{
    const double inv_csx = 1.0 / g->cell_size_x;
    const double inv_csy = 1.0 / g->cell_size_y;
    const double fcx = (wx - g->min_x) * inv_csx;
    const double fcy = (wy - g->max_y) * inv_csy;
    const double flx = floor(fcx), fly = floor(fcy);
    const float subx = (float)(fcx - flx);
    const float suby = (float)(fcy - fly);

    const float sxw = (rd->sigma_x && rd->sigma_x[p] > 0.0f) ? rd->sigma_x[p]
                                                             : rd->default_sigma_x;
    const float syw = (rd->sigma_y && rd->sigma_y[p] > 0.0f) ? rd->sigma_y[p]
                                                             : rd->default_sigma_y;
    const float sx = sxw * (float)inv_csx;
    const float sy = syw * (float)inv_csy;          /* negative for north-up */

    const float rot = rd->rotation ? rd->rotation[p] : rd->default_rotation;
    const float cr = cosf(-rot);
    const float sr = sinf(-rot);

    /* with sy < 0 the radius depends on sx only */
    const float R = std_minf(3.0f * std_maxf(sx, sy), rd->max_radius_cells);
    const int32_t r = (int32_t)ceilf(R);
    const int32_t icx = (int32_t)flx, icy = (int32_t)fly;

    for (int32_t dy = -r; dy <= r; ++dy) {
        for (int32_t dx = -r; dx <= r; ++dx) {
            const int32_t gc = icx + dx, gr = icy + dy;
            const int32_t lc = gc - t.c0, lr = gr - t.r0;
            if (lc < 0 || lc >= t.w || lr < 0 || lr >= t.h) continue;
            /* weight sampled at the cell's integer corner, not its centre */
            const float ox = (float)dx - subx;
            const float oy = (float)dy - suby;
            const float rx = ox * cr + oy * (-sr);
            const float ry = ox * sr + oy * cr;
            const float qx = rx / sx, qy = ry / sy;
            const float w = expf(-0.5f * (qx * qx + qy * qy));
            if (w < 1e-6f) continue;
            visit(ctx, (int64_t)gr * g->width + gc, w);
        }
    }
}

/* ------------------------------------------------------------------------ */
/* Ingest: Pipeline::Impl::process_cloud, src/engine/pipeline.cpp:283-770     */
/* ------------------------------------------------------------------------ */

typedef struct { int type; float *state; int64_t cells; float v; } fold_ctx;

static void fold_visit(void *c, int64_t cell, float w)
{
    fold_ctx *f = (fold_ctx *)c;
    combine_glyph(f->type, f->state, f->cells, cell, f->v, w);
}

/* Reject what the reference rejects with NotImplemented
 * (pipeline.cpp:500-508, glyph_kernels.cu:296-302). */
static int glyph_combo_ok(const orc_reduction *rd)
{
    return rd->glyph == ORC_GLYPH_POINT ||
           !(rd->type == ORC_MAX || rd->type == ORC_MIN);
}

/* Accumulates one cloud into one reduction's full-grid state.  Points are
 * folded in input order (the reference folds them in (tile,cell)-sorted order
 * from an unstable std::sort, tile_router.cpp:159-172, so its intra-cell order
 * is unspecified; Count/Max/Min are order-free, float sums are compared under
 * a tolerance).  touched[] gets 1 for every tile that received a batch.
 * Returns 0, or -1 for a rejected glyph/reduction combination. */
int orc_accumulate(const orc_grid *g, const orc_reduction *rd,
                   const double *x, const double *y, size_t n,
                   float *state, uint8_t *touched)
{
    const int64_t cells = (int64_t)g->width * g->height;
    const int32_t tx = tiles_x_of(g);
    if (!glyph_combo_ok(rd)) return -1;

    for (size_t p = 0; p < n; ++p) {
        int32_t col, row;
        if (!orc_world_to_cell(g, x[p], y[p], &col, &row)) continue;
        touched[(row / g->tile_height) * tx + col / g->tile_width] = 1;
        if (rd->glyph == ORC_GLYPH_POINT) {
            combine_point(rd->type, state, cells, (int64_t)row * g->width + col,
                          rd->value[p]);
            continue;
        }
        fold_ctx f = { rd->type, state, cells, rd->value[p] };
        if (rd->glyph == ORC_GLYPH_LINE)
            footprint_line(g, rd, p, x[p], y[p], tile_of(g, col, row),
                           fold_visit, &f);
        else
            footprint_gaussian(g, rd, p, x[p], y[p], tile_of(g, col, row),
                               fold_visit, &f);
    }
    return 0;
}

/* Error-bound helper for the float-sum parity tests (not in the reference).
 * Per cell: the f64 sum of contributions, the f64 sum of |contribution| and the
 * number of contributions — of the value plane (v, or v*w for glyphs) when
 * want_weight == 0, of the weight plane (1, or w) otherwise.  The footprints
 * come from the same enumerators orc_accumulate uses. */
typedef struct { double *sum, *abs; uint32_t *cnt; float v; int want_weight; }
    bound_ctx;

static void bound_visit(void *c, int64_t cell, float w)
{
    bound_ctx *b = (bound_ctx *)c;
    const double t = b->want_weight ? (double)w : (double)(b->v * w);
    b->sum[cell] += t;
    b->abs[cell] += fabs(t);
    b->cnt[cell] += 1;
}

int orc_accumulate_bounds(const orc_grid *g, const orc_reduction *rd,
                          const double *x, const double *y, size_t n,
                          int want_weight, double *sum64, double *abs64,
                          uint32_t *cnt)
{
    if (!glyph_combo_ok(rd)) return -1;
    for (size_t p = 0; p < n; ++p) {
        int32_t col, row;
        if (!orc_world_to_cell(g, x[p], y[p], &col, &row)) continue;
        bound_ctx b = { sum64, abs64, cnt, rd->value[p], want_weight };
        if (rd->glyph == ORC_GLYPH_POINT)
            bound_visit(&b, (int64_t)row * g->width + col, 1.0f);
        else if (rd->glyph == ORC_GLYPH_LINE)
            footprint_line(g, rd, p, x[p], y[p], tile_of(g, col, row),
                           bound_visit, &b);
        else
            footprint_gaussian(g, rd, p, x[p], y[p], tile_of(g, col, row),
                               bound_visit, &b);
    }
    return 0;
}
