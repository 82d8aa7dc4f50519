// emosaic.cpp — see emosaic.hpp.  Host glue only; every pixel/index comes out of libemosaic_cuda.so.
#include "emosaic.hpp"

#include <tuple>

#include <queue>

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <random>
#include <sys/stat.h>

namespace emosaic {

static uint32_t isqrt_exact(uint32_t N) {
    uint32_t d = (uint32_t)std::floor(std::sqrt((double)N));
    while ((uint64_t)d * d > N) d--;
    while ((uint64_t)(d + 1) * (d + 1) <= N) d++;
    if (N == 0 || d * d != N) throw Error(EMO_ERR_ARG, "N=" + std::to_string(N) + " is not a square number of cells");
    return d;
}

void check(int rc) {
    if (rc != 0) throw Error(rc, emo_last_error());
}

Context::Context(int device) { check(emo_create(device, &h_)); }
Context::Context(const std::vector<int> &devices) {
    if (devices.size() <= 1) {
        check(emo_create(devices.empty() ? 0 : devices[0], &h_));
        return;
    }
    check(emo_group_create(devices.data(), (int)devices.size(), &g_));
    h_ = emo_group_ctx(g_, 0);
}
Context::~Context() {
    if (g_) emo_group_destroy(g_);  // owns its members
    else emo_destroy(h_);
}
void Context::build_index() const {
    for (int i = 0; i < gpus(); i++) check(emo_build_index(g_ ? emo_group_ctx(g_, i) : h_));
}
void Context::set_match_mode(int mode) const {
    for (int i = 0; i < gpus(); i++) check(emo_set_match_mode(g_ ? emo_group_ctx(g_, i) : h_, mode));
}

// ---- tiles/utils.rs:18-43 ------------------------------------------------------------------------
void flipped_coords(std::vector<uint32_t> &coords) {
    const size_t n = coords.size();
    size_t rows = (size_t)std::floor(std::sqrt((double)(n / 3)));
    while (rows * rows > n / 3) rows--;
    const size_t cols = rows, cir = cols * 3;
    for (size_t i = 0; i < rows; i++)
        for (size_t j = 0; j < cols / 2; j++) {
            const size_t a = i * cir + j * 3, b = (i + 1) * cir - (j + 1) * 3;
            for (size_t h = 0; h < 3; h++) std::swap(coords[a + h], coords[b + h]);
        }
}

std::vector<uint32_t> Tile::coords() const {
    std::vector<uint32_t> r(colors.begin(), colors.end());
    if (flipped) flipped_coords(r);
    return r;
}

// ---- analysis.rs ----------------------------------------------------------------------------------
std::vector<uint8_t> analyse_tiles(Context &ctx, const std::vector<Image> &tiles, uint32_t N) {
    const uint32_t dim = isqrt_exact(N);
    if (tiles.empty()) return {};
    const uint32_t ts = tiles[0].width;
    std::vector<uint8_t> px((size_t)tiles.size() * ts * ts * 3);
    for (size_t t = 0; t < tiles.size(); t++) {
        if (tiles[t].width != ts || tiles[t].height != ts || tiles[t].channels != 3)
            throw Error(EMO_ERR_ARG, "analyse: tiles are square RGB images of one size on this path");
        std::memcpy(px.data() + t * (size_t)ts * ts * 3, tiles[t].data.data(), (size_t)ts * ts * 3);
    }
    std::vector<uint8_t> out((size_t)tiles.size() * N * 3);
    if (ctx.group()) check(emo_group_analyse(ctx.group(), px.data(), tiles.size(), ts, dim, out.data()));
    else check(emo_analyse(ctx.handle(), px.data(), tiles.size(), ts, dim, out.data()));
    return out;
}

std::vector<uint8_t> analyse(Context &ctx, const Image &img, uint32_t N) { return analyse_tiles(ctx, {img}, N); }

std::vector<uint8_t> get_img_colors(uint32_t x, uint32_t y, uint32_t step, const Image &src, uint32_t N) {
    std::vector<uint8_t> out((size_t)N * 3);
    for (uint32_t i = 0; i < N; i++) std::memcpy(&out[3 * i], src.pixel(x + i % step, y + i / step), 3);
    return out;
}

// ---- tiles/tileset.rs -----------------------------------------------------------------------------
void TileSet::push_tile(const std::string &path, std::vector<uint8_t> colors, std::optional<std::string> date) {
    if (colors.size() != (size_t)N_ * 3) throw Error(EMO_ERR_ARG, "push_tile: colours must hold N*3 bytes");
    const uint16_t idx = (uint16_t)(tiles_.size() + 1);  // tileset.rs:112 `len() as u16 + 1`
    tiles_.push_back(Tile{std::move(colors), idx, false, std::move(date)});
    paths_.push_back(path);
    images_.emplace_back();
}

void TileSet::push_tile_with_image(const std::string &path, std::vector<uint8_t> colors, Image image) {
    push_tile(path, std::move(colors));
    images_.back() = std::move(image);
}

std::optional<Tile> TileSet::get_tile(int32_t idx) const {
    const size_t a = (size_t)(idx < 0 ? -(int64_t)idx : idx);
    if (a == 0 || a > tiles_.size()) return std::nullopt;
    Tile t = tiles_[a - 1];
    t.flipped = idx < 0;
    return t;
}

Image TileSet::get_image(const Tile &tile, uint32_t tile_size) const {
    const Image &im = images_.at(tile.idx - 1);
    if (im.data.empty() || im.width != tile_size || im.height != tile_size)
        throw Error(EMO_ERR_ARG, "Image not found: " + paths_.at(tile.idx - 1) + " (prepare_tile is outside the accelerated path)");
    if (!tile.flipped) return im;
    Image f(im.width, im.height, 3);  // imageops::flip_horizontal
    for (uint32_t y = 0; y < im.height; y++)
        for (uint32_t x = 0; x < im.width; x++) std::memcpy(f.pixel(x, y), im.pixel(im.width - 1 - x, y), 3);
    return f;
}

void TileSet::build_kiddo(Context &ctx, uint32_t tile_size) const {
    if (tiles_.empty()) throw Error(EMO_ERR_ARG, "empty tile set");
    std::vector<uint8_t> colors((size_t)tiles_.size() * N_ * 3), px;
    for (size_t t = 0; t < tiles_.size(); t++) std::memcpy(&colors[t * (size_t)N_ * 3], tiles_[t].colors.data(), (size_t)N_ * 3);
    if (tile_size) {
        px.resize((size_t)tiles_.size() * tile_size * tile_size * 3);
        for (size_t t = 0; t < tiles_.size(); t++) {
            const Image &im = images_[t];
            if (im.data.empty() || im.width != tile_size || im.height != tile_size || im.channels != 3)
                throw Error(EMO_ERR_ARG, "Image not found: " + paths_[t] + " (tiles must be prepared on the host)");
            std::memcpy(&px[t * (size_t)tile_size * tile_size * 3], im.data.data(), (size_t)tile_size * tile_size * 3);
        }
    }
    if (ctx.group())
        check(emo_group_set_library(ctx.group(), colors.data(), tile_size ? px.data() : nullptr, (uint32_t)tiles_.size(), N_, tile_size));
    else
        check(emo_set_library(ctx.handle(), colors.data(), tile_size ? px.data() : nullptr, (uint32_t)tiles_.size(), N_, tile_size));
}

// ---- rendering.rs ---------------------------------------------------------------------------------
uint8_t tint_alpha(double t) {
    const double v = 255.0 * t;
    if (!(v > 0.0)) return 0;
    return v >= 255.0 ? 255 : (uint8_t)v;  // Rust `as u8`: saturating, truncating
}

RenderResult render_nto1(Context &ctx, const Image &source, const TileSet &tile_set, uint32_t tile_size, bool no_repeat,
                         std::optional<double> randomize, double tint_opacity) {
    if (no_repeat || randomize)
        throw Error(EMO_ERR_UNSUPPORTED,
                    "no_repeat (greedy, rayon-order dependent) / randomize (thread_rng) inside render_nto1 are outside the accelerated "
                    "path; the deterministic no-repeat renderer is render_nto1_no_repeat");
    const uint32_t dim = isqrt_exact(tile_set.cells());
    if (source.channels != 3) throw Error(EMO_ERR_ARG, "render_nto1: RGB source expected");
    if (source.width % dim || source.height % dim)  // main.rs:603-611
        throw Error(EMO_ERR_ARG, "Invalid source dimensions (" + std::to_string(source.width) + "x" + std::to_string(source.height) +
                                     "): Dimensions must be divisible by " + std::to_string(dim));
    if (tile_size % dim) throw Error(EMO_ERR_ARG, "Invalid tile size: Tile size must be divisible by " + std::to_string(dim));
    tile_set.build_kiddo(ctx, tile_size);
    RenderResult r;
    r.bw = source.width / dim;
    r.bh = source.height / dim;
    const uint32_t oc = tint_opacity > 0.0 ? 4 : 3;
    r.image = Image(r.bw * tile_size, r.bh * tile_size, oc);
    r.item.resize((size_t)r.bw * r.bh);
    r.dist.resize((size_t)r.bw * r.bh);
    if (ctx.group())  // row stripes over the GPUs, each copied to its offset of r.image (rendering.rs:68-101)
        check(emo_group_mosaic(ctx.group(), source.data.data(), source.width, source.height, oc, tint_alpha(tint_opacity), r.item.data(),
                               r.dist.data(), r.image.data.data()));
    else
        check(emo_mosaic(ctx.handle(), source.data.data(), source.width, source.height, oc, tint_alpha(tint_opacity), r.item.data(),
                         r.dist.data(), r.image.data.data()));
    return r;
}

// rendering.rs:262-401.  The ranked candidate lists (Scoring phase, :307-321) come from the GPU in pages (emo_topk); the
// greedy merge (:341-392) runs here: blocks ordered by the distance of their best remaining candidate, ties by the
// reference's block number n = bx * vtiles + by (:300-301; the canonical order), a block whose candidate is taken
// moves on to its next one, a block that runs out of candidates stays black (:347-351).
RenderResult render_nto1_no_repeat(Context &ctx, const Image &source, const TileSet &tile_set, uint32_t tile_size, uint32_t page) {
    const uint32_t dim = isqrt_exact(tile_set.cells());
    if (source.channels != 3) throw Error(EMO_ERR_ARG, "render_nto1_no_repeat: RGB source expected");
    if (source.width % dim || source.height % dim)
        throw Error(EMO_ERR_ARG, "Invalid source dimensions (" + std::to_string(source.width) + "x" + std::to_string(source.height) +
                                     "): Dimensions must be divisible by " + std::to_string(dim));
    if (tile_size % dim) throw Error(EMO_ERR_ARG, "Invalid tile size: Tile size must be divisible by " + std::to_string(dim));
    const uint32_t bw = source.width / dim, bh = source.height / dim;
    const size_t Q = (size_t)bw * bh, T = tile_set.len();
    if (Q > 2 * T)  // rendering.rs:292-298
        throw Error(EMO_ERR_ARG, "Insufficient tiles for no-repeat mode: need " + std::to_string(Q) + " tiles but only have " +
                                     std::to_string(2 * T) + " available");
    tile_set.build_kiddo(ctx, tile_size);
    RenderResult r;
    r.bw = bw;
    r.bh = bh;
    r.item.assign(Q, 0);
    r.dist.assign(Q, 0);
    // scoring (:307-321), ordering (:323-326) and the greedy loop (:341-392) in one library call
    check(emo_no_repeat(ctx.handle(), source.data.data(), source.width, source.height, page, r.item.data(), r.dist.data(), nullptr));
    std::vector<int32_t> placed(r.item);
    bool holes = false;
    for (auto &v : placed)
        if (v == 0) {
            v = 1;
            holes = true;
        }
    r.image = Image(bw * tile_size, bh * tile_size, 3);
    check(emo_compose(ctx.handle(), placed.data(), nullptr, source.width, source.height, 3, 0, r.image.data.data()));
    if (holes)  // RgbImage::new: unplaced blocks stay black
        for (size_t q = 0; q < Q; q++)
            if (r.item[q] == 0)
                for (uint32_t y = 0; y < tile_size; y++)
                    std::memset(r.image.pixel((uint32_t)(q % bw) * tile_size, (uint32_t)(q / bw) * tile_size + y), 0, (size_t)tile_size * 3);
    return r;
}

Image render_random(Context &ctx, const Image &source, const TileSet &tile_set, uint32_t tile_size, uint64_t seed) {
    // rendering.rs:418-440: a uniformly random tile per source pixel; only the composition is accelerated
    std::vector<uint8_t> px((size_t)tile_set.len() * tile_size * tile_size * 3), colors((size_t)tile_set.len() * 3, 0);
    for (size_t t = 0; t < tile_set.len(); t++) {
        Image im = tile_set.get_image(*tile_set.get_tile((int32_t)t + 1), tile_size);
        std::memcpy(&px[t * (size_t)tile_size * tile_size * 3], im.data.data(), im.data.size());
    }
    check(emo_set_library(ctx.handle(), colors.data(), px.data(), (uint32_t)tile_set.len(), 1, tile_size));
    std::mt19937_64 rng(seed);
    std::vector<int32_t> item((size_t)source.width * source.height);
    for (auto &v : item) v = (int32_t)(rng() % tile_set.len()) + 1;
    Image out(source.width * tile_size, source.height * tile_size, 3);
    check(emo_compose(ctx.handle(), item.data(), nullptr, source.width, source.height, 3, 0, out.data.data()));
    return out;
}

// ---- main.rs:567-601 --------------------------------------------------------------------------------
std::pair<uint32_t, uint32_t> adjust_dims(uint32_t w, uint32_t h, uint32_t downsample, uint32_t dim) {
    uint32_t nw = w / downsample, nh = h / downsample, m = nw % dim;
    nw = m > dim / 2 ? nw + dim - m : nw - m;
    m = nh % dim;
    nh = m > dim / 2 ? nh + dim - m : nh - m;
    return {nw, nh};
}

// ---- resize / prepare_tile ----------------------------------------------------------------------------
Image resize_lanczos3(Context &ctx, const Image &img, uint32_t nw, uint32_t nh, std::optional<View> view) {
    if (img.channels != 3) throw Error(EMO_ERR_ARG, "resize: RGB images only");
    const View v = view.value_or(View{0, 0, img.width, img.height});
    Image out(nw, nh, 3);
    check(emo_resize(ctx.handle(), img.data.data(), 1, img.width, img.height, v.x, v.y, v.w, v.h, nw, nh, out.data.data()));
    return out;
}

Image resize_source(Context &ctx, const Image &original, uint32_t downsample, uint32_t dim) {
    auto [nw, nh] = adjust_dims(original.width, original.height, downsample, dim);
    fprintf(stderr, "Resizing source image from %ux%u to %ux%u\n", original.width, original.height, nw, nh);  // main.rs:587-593
    return resize_lanczos3(ctx, original, nw, nh);
}

uint32_t most_common_value(const std::vector<uint32_t> &values) {
    std::map<uint32_t, uint32_t> counts;  // ordered: the first maximum is the smallest value (canonical tie rule)
    for (uint32_t v : values) counts[v]++;
    uint32_t best = 0, bc = 0;            // empty input: unwrap_or((0, 0)).0
    for (auto &[v, c] : counts)
        if (c > bc) { bc = c; best = v; }
    return best;
}

View prepare_view(const Image &img, uint32_t tile_size, bool crop) {
    const uint32_t w = img.width, h = img.height;
    if (w < tile_size || h < tile_size)  // utils.rs:99-106
        throw Error(EMO_ERR_ARG, "Image " + std::to_string(w) + "x" + std::to_string(h) + " is smaller than the tile size (DimensionError)");
    auto white = [&](uint32_t x, uint32_t y) { const uint8_t *p = img.pixel(x, y); return p[0] > 240 && p[1] > 240 && p[2] > 240; };
    std::vector<uint32_t> from_left, from_right, from_top, from_bottom;
    for (uint32_t y = 0; y < h; y++) {
        uint32_t x = 0;
        while (x < w && white(x, y)) x++;
        if (x != w) from_left.push_back(x);
        uint32_t r = 0;
        for (uint32_t k = w; k > x; k--)
            if (!white(k - 1, y)) { r = k - 1; break; }
        if (r != 0) from_right.push_back(r);
    }
    for (uint32_t x = 0; x < w; x++) {
        uint32_t y = 0;
        while (y < h && white(x, y)) y++;
        if (y != h) from_top.push_back(y);
        uint32_t r = 0;
        for (uint32_t k = h; k > y; k--)
            if (!white(x, k - 1)) { r = k - 1; break; }
        if (r != 0) from_bottom.push_back(r);
    }
    const uint32_t c0 = most_common_value(from_left), c1 = most_common_value(from_right), r0 = most_common_value(from_top),
                   r1 = most_common_value(from_bottom);
    if (!(c0 < c1)) throw Error(EMO_ERR_ARG, "assertion failed: first_non_white_col < last_non_white_col");  // utils.rs:157
    if (!(r0 < r1)) throw Error(EMO_ERR_ARG, "assertion failed: first_non_white_row < last_non_white_row");  // utils.rs:158
    View v{c0, r0, c1 - c0, r1 - r0};
    if (crop) {  // utils.rs:170-182
        const uint32_t size = std::min(v.w, v.h);
        v = View{v.x + (v.w - size) / 2, v.y + (v.h - size) / 2, size, size};
    }
    return v;
}

Image rotate(const Image &img, uint32_t orientation) {
    if (orientation < 2 || orientation > 8) return img;
    const bool swap = orientation >= 5;
    Image out(swap ? img.height : img.width, swap ? img.width : img.height, img.channels);
    for (uint32_t y = 0; y < out.height; y++)
        for (uint32_t x = 0; x < out.width; x++) {
            uint32_t sx = x, sy = y;
            switch (orientation) {
                case 2: sx = img.width - 1 - x; break;                              // flip_horizontal
                case 3: sx = img.width - 1 - x; sy = img.height - 1 - y; break;     // rotate180
                case 4: sy = img.height - 1 - y; break;                             // flip_vertical
                case 5: sx = y; sy = x; break;                                      // flip_horizontal(rotate90) = transpose
                case 6: sx = y; sy = img.height - 1 - x; break;                     // rotate90 (clockwise)
                case 7: sx = img.width - 1 - y; sy = img.height - 1 - x; break;     // flip_horizontal(rotate270) = anti-transpose
                case 8: sx = img.width - 1 - y; sy = x; break;                      // rotate270
            }
            std::memcpy(out.pixel(x, y), img.pixel(sx, sy), img.channels);
        }
    return out;
}

Image prepare_tile(Context &ctx, const Image &decoded, uint32_t tile_size, bool crop, uint32_t orientation) {
    return rotate(resize_lanczos3(ctx, decoded, tile_size, tile_size, prepare_view(decoded, tile_size, crop)), orientation);
}

std::string cache_file_name(uint32_t N, bool crop) { return ".emosaic_" + std::to_string(N) + "to1" + (crop ? "_cropped" : ""); }

// ---- cache ------------------------------------------------------------------------------------------
static void put_u64(std::vector<uint8_t> &o, uint64_t v) {
    for (int i = 0; i < 8; i++) o.push_back((uint8_t)(v >> (8 * i)));
}

std::vector<uint8_t> serialize_tile_set(const TileSet &ts) {
    std::vector<uint8_t> o;
    put_u64(o, ts.len());
    for (const Tile &t : ts.tiles()) {
        put_u64(o, t.colors.size());
        o.insert(o.end(), t.colors.begin(), t.colors.end());
        o.push_back((uint8_t)t.idx);
        o.push_back((uint8_t)(t.idx >> 8));
        if (t.date_taken) {
            o.push_back(1);
            put_u64(o, t.date_taken->size());
            o.insert(o.end(), t.date_taken->begin(), t.date_taken->end());
        } else {
            o.push_back(0);
        }
    }
    put_u64(o, ts.len());
    for (const std::string &p : ts.paths()) {
        put_u64(o, p.size());
        o.insert(o.end(), p.begin(), p.end());
    }
    return o;
}

TileSet deserialize_tile_set(const std::vector<uint8_t> &b, uint32_t N, const std::vector<std::string> *extensions, bool check_exists) {
    size_t pos = 0;
    auto need = [&](uint64_t n) {  // n comes from untrusted u64 lengths: compare without forming pos + n
        if (n > b.size() - pos) throw Error(EMO_ERR_ARG, "truncated cache file");
    };
    auto get_u64 = [&]() {
        need(8);
        uint64_t v = 0;
        for (int i = 0; i < 8; i++) v |= (uint64_t)b[pos + i] << (8 * i);
        pos += 8;
        return v;
    };
    const uint64_t T = get_u64();
    // every tile costs at least 8 + 3N + 3 bytes here and 8 more in the path list: a corrupt count fails before any allocation
    if (T > (b.size() - pos) / ((uint64_t)N * 3 + 19)) throw Error(EMO_ERR_ARG, "truncated cache file");
    std::vector<std::vector<uint8_t>> colors;
    std::vector<std::optional<std::string>> dates;
    for (uint64_t t = 0; t < T; t++) {
        const uint64_t ln = get_u64();
        if (ln != (uint64_t)N * 3) throw Error(EMO_ERR_ARG, "cache entry has the wrong vector length");  // try_into().unwrap()
        need(ln + 3);  // ln == 3N was checked above, no overflow
        colors.emplace_back(b.begin() + pos, b.begin() + pos + ln);
        pos += ln + 2;  // stored idx ignored: renumbered below (main.rs:643-652)
        const uint8_t tag = b[pos++];
        if (tag == 1) {
            const uint64_t dl = get_u64();
            need(dl);
            dates.emplace_back(std::string(b.begin() + pos, b.begin() + pos + dl));
            pos += dl;
        } else if (tag == 0) {
            dates.emplace_back(std::nullopt);
        } else {
            throw Error(EMO_ERR_ARG, "bad Option tag in cache file");
        }
    }
    if (get_u64() != T) throw Error(EMO_ERR_ARG, "tiles / paths length mismatch in cache file");
    TileSet ts(N);
    for (uint64_t t = 0; t < T; t++) {
        const uint64_t pl = get_u64();
        need(pl);
        std::string p(b.begin() + pos, b.begin() + pos + pl);
        pos += pl;
        bool keep = true;
        if (extensions || check_exists) {
            const size_t dot = p.find_last_of('.'), slash = p.find_last_of('/');
            const std::string ext = (dot == std::string::npos || (slash != std::string::npos && dot < slash)) ? "" : p.substr(dot + 1);
            if (ext.empty()) keep = false;
            if (keep && extensions && std::find(extensions->begin(), extensions->end(), ext) == extensions->end()) keep = false;
            struct stat st;
            if (keep && check_exists && stat(p.c_str(), &st) != 0) keep = false;
        }
        if (keep) ts.push_tile(p, colors[t], dates[t]);
    }
    return ts;
}

// ---- stats.rs ---------------------------------------------------------------------------------------
StatsSummary summarise(const RenderResult &r, const TileSet &ts, bool print, Context *ctx) {
    StatsSummary s;
    // rendering.rs:352-365 records placed tiles only: a no-repeat block that ran out of tiles (item 0) has no entry
    std::vector<size_t> placed;
    for (size_t i = 0; i < r.item.size(); i++)
        if (r.item[i] != 0) placed.push_back(i);
    s.total = placed.size();
    if (!s.total) {
        if (print) fprintf(stderr, "No tiles recorded in statistics\n");
        return s;
    }
    std::map<uint32_t, uint32_t> usage;
    uint64_t sum = 0;
    if (ctx) {  // one pass over the maps on the GPU (stats.cu)
        uint64_t sums[3] = {0, 0, 0};
        std::vector<uint32_t> u(ts.len(), 0);
        check(emo_stats(ctx->handle(), r.item.data(), r.dist.data(), r.item.size(), (uint32_t)ts.len(), sums, u.data()));
        for (size_t t = 0; t < u.size(); t++)
            if (u[t]) usage[(uint32_t)t + 1] = u[t];
        sum = sums[1];
        s.total = (size_t)sums[0];
    } else {
        for (size_t i : placed) {
            usage[(uint32_t)std::abs(r.item[i])]++;
            sum += r.dist[i];
        }
    }
    s.unique = usage.size();
    s.average_distance = (double)sum / (double)s.total;
    std::vector<std::pair<uint32_t, uint32_t>> u(usage.begin(), usage.end());
    std::stable_sort(u.begin(), u.end(), [](auto &a, auto &b) { return a.second > b.second; });
    for (size_t i = 0; i < u.size() && i < 10; i++) s.top.emplace_back(ts.paths().at(u[i].first - 1), u[i].second);
    std::vector<size_t> order = placed;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return r.dist[a] > r.dist[b]; });
    for (size_t i = 0; i < order.size() && i < 10; i++)
        s.worst.emplace_back(ts.paths().at((size_t)std::abs(r.item[order[i]]) - 1), r.dist[order[i]]);
    if (print) {
        fprintf(stderr, "Mosaic Statistics:\n  Total tiles placed: %zu\n  Unique images used: %zu\n  Average color distance: %.3f\n",
                s.total, s.unique, s.average_distance);
        fprintf(stderr, "\nTop 10 most used tiles:\n");
        for (size_t i = 0; i < s.top.size(); i++) fprintf(stderr, "  %zu. %s (%u times)\n", i + 1, s.top[i].first.c_str(), s.top[i].second);
        fprintf(stderr, "\nWorst 10 color matches:\n");
        for (size_t i = 0; i < s.worst.size(); i++)
            fprintf(stderr, "  %zu. %s (distance: %u)\n", i + 1, s.worst[i].first.c_str(), s.worst[i].second);
    }
    return s;
}

Image render_stats(const RenderResult &r, uint32_t dim, uint32_t tile_size) {
    // `dim` = distance between the recorded coordinates of neighbouring blocks: render_nto1 records SOURCE coordinates
    // (step = dim, rendering.rs:211-214), render_nto1_no_repeat OUTPUT coordinates (step = tile_size, rendering.rs:352-365)
    bool any = false;
    for (int32_t it : r.item) any = any || it != 0;
    if (!any) throw Error(EMO_ERR_ARG, "Cannot render visualization: no tiles recorded");
    if (tile_size == 0) throw Error(EMO_ERR_ARG, "Tile size must be greater than 0");
    uint32_t max_x = 0, max_y = 0, md = 0;
    for (uint32_t by = 0; by < r.bh; by++)
        for (uint32_t bx = 0; bx < r.bw; bx++)
            if (r.item[(size_t)by * r.bw + bx] != 0) {
                max_x = std::max(max_x, bx * dim);
                max_y = std::max(max_y, by * dim);
                md = std::max(md, r.dist[(size_t)by * r.bw + bx]);
            }
    Image img(max_x / tile_size + 1, max_y / tile_size + 1, 3);
    for (uint32_t by = 0; by < r.bh; by++)
        for (uint32_t bx = 0; bx < r.bw; bx++) {
            if (r.item[(size_t)by * r.bw + bx] == 0) continue;
            const double nd = md > 0 ? (double)r.dist[(size_t)by * r.bw + bx] / (double)md : 0.0;
            const uint8_t v = (uint8_t)(nd * 255.0);
            uint8_t *p = img.pixel(bx * dim / tile_size, by * dim / tile_size);  // stats.rs:176-191 (source coords / tile_size)
            p[0] = p[1] = p[2] = v;
        }
    return img;
}

// ---- image I/O ----------------------------------------------------------------------------------------
static std::vector<uint8_t> slurp(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error(EMO_ERR_ARG, "Failed to open " + path);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

static Image read_ppm(const std::vector<uint8_t> &b, const std::string &path) {
    size_t pos = 2;
    auto token = [&]() {
        while (pos < b.size() && (isspace(b[pos]) || b[pos] == '#')) {
            if (b[pos] == '#') while (pos < b.size() && b[pos] != '\n') pos++;
            else pos++;
        }
        size_t s = pos;
        while (pos < b.size() && !isspace(b[pos])) pos++;
        return std::stoul(std::string(b.begin() + s, b.begin() + pos));
    };
    const uint32_t w = token(), h = token(), mx = token();
    pos++;
    if (mx != 255 || pos + (size_t)w * h * 3 > b.size()) throw Error(EMO_ERR_ARG, "unsupported PPM: " + path);
    Image im(w, h, 3);
    std::memcpy(im.data.data(), b.data() + pos, im.data.size());
    return im;
}

static uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

static Image read_png(const std::vector<uint8_t> &b, const std::string &path) {
    size_t pos = 8;
    uint32_t w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    while (pos + 12 <= b.size()) {
        const uint32_t len = be32(&b[pos]);
        const std::string type(b.begin() + pos + 4, b.begin() + pos + 8);
        const uint8_t *d = &b[pos + 8];
        if (pos + 12 + len > b.size()) break;
        if (type == "IHDR") { w = be32(d); h = be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12]; }
        else if (type == "PLTE") plte.assign(d, d + len);
        else if (type == "IDAT") idat.insert(idat.end(), d, d + len);
        else if (type == "IEND") break;
        pos += 12 + len;
    }
    const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!w || !h || depth != 8 || !ch || interlace) throw Error(EMO_ERR_ARG, "unsupported PNG (need 8-bit, non-interlaced): " + path);
    const size_t stride = (size_t)w * ch;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rl = raw.size();
    if (uncompress(raw.data(), &rl, idat.data(), idat.size()) != Z_OK || rl != raw.size()) throw Error(EMO_ERR_ARG, "corrupt PNG: " + path);
    std::vector<uint8_t> px(stride * h);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t ft = raw[y * (stride + 1)];
        const uint8_t *s = &raw[y * (stride + 1) + 1];
        uint8_t *o = &px[y * stride];
        const uint8_t *up = y ? &px[(y - 1) * stride] : nullptr;
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= (size_t)ch ? o[i - ch] : 0, bb = up ? up[i] : 0, c = (up && i >= (size_t)ch) ? up[i - ch] : 0;
            int pred = 0;
            if (ft == 1) pred = a;
            else if (ft == 2) pred = bb;
            else if (ft == 3) pred = (a + bb) / 2;
            else if (ft == 4) {
                const int p = a + bb - c, pa = std::abs(p - a), pb = std::abs(p - bb), pc = std::abs(p - c);
                pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? bb : c);
            }
            o[i] = (uint8_t)(s[i] + pred);
        }
    }
    Image im(w, h, 3);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        uint8_t *o = &im.data[i * 3];
        const uint8_t *s = &px[i * ch];
        if (ctype == 2 || ctype == 6) { o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; }
        else if (ctype == 0 || ctype == 4) { o[0] = o[1] = o[2] = s[0]; }
        else { if ((size_t)s[0] * 3 + 2 >= plte.size()) throw Error(EMO_ERR_ARG, "bad palette index: " + path);
               o[0] = plte[s[0] * 3]; o[1] = plte[s[0] * 3 + 1]; o[2] = plte[s[0] * 3 + 2]; }
    }
    return im;
}

Image read_image(const std::string &path) {
    const std::vector<uint8_t> b = slurp(path);
    if (b.size() > 8 && !std::memcmp(b.data(), "\x89PNG\r\n\x1a\n", 8)) return read_png(b, path);
    if (b.size() > 2 && b[0] == 'P' && b[1] == '6') return read_ppm(b, path);
    throw Error(EMO_ERR_UNSUPPORTED, "only PNG (8-bit, non-interlaced) and PPM (P6) are decoded by the C++ host; JPEG decoding stays in "
                                     "the Python front end (python -m emosaic_b200): " + path);
}

void write_ppm(const std::string &path, const Image &img) {
    std::ofstream f(path, std::ios::binary);
    f << "P6\n" << img.width << " " << img.height << "\n255\n";
    if (img.channels == 3) f.write((const char *)img.data.data(), img.data.size());
    else for (size_t i = 0; i < (size_t)img.width * img.height; i++) f.write((const char *)&img.data[i * img.channels], 3);
}

static void put_chunk(std::vector<uint8_t> &o, const char *type, const uint8_t *d, size_t n) {
    const uint8_t l[4] = {(uint8_t)(n >> 24), (uint8_t)(n >> 16), (uint8_t)(n >> 8), (uint8_t)n};
    o.insert(o.end(), l, l + 4);
    const size_t s = o.size();
    o.insert(o.end(), type, type + 4);
    if (n) o.insert(o.end(), d, d + n);
    const uint32_t c = crc32(0, &o[s], (uInt)(n + 4));
    const uint8_t cc[4] = {(uint8_t)(c >> 24), (uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c};
    o.insert(o.end(), cc, cc + 4);
}

void write_png(const std::string &path, const Image &img) {
    const size_t stride = (size_t)img.width * img.channels;
    std::vector<uint8_t> raw((stride + 1) * img.height);
    for (uint32_t y = 0; y < img.height; y++) {
        raw[y * (stride + 1)] = 0;
        std::memcpy(&raw[y * (stride + 1) + 1], &img.data[y * stride], stride);
    }
    uLongf cl = compressBound(raw.size());
    std::vector<uint8_t> z(cl);
    if (compress2(z.data(), &cl, raw.data(), raw.size(), 1) != Z_OK) throw Error(EMO_ERR_ARG, "PNG compression failed");
    std::vector<uint8_t> o = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    uint8_t ihdr[13] = {(uint8_t)(img.width >> 24), (uint8_t)(img.width >> 16), (uint8_t)(img.width >> 8), (uint8_t)img.width,
                        (uint8_t)(img.height >> 24), (uint8_t)(img.height >> 16), (uint8_t)(img.height >> 8), (uint8_t)img.height,
                        8, (uint8_t)(img.channels == 4 ? 6 : 2), 0, 0, 0};
    put_chunk(o, "IHDR", ihdr, 13);
    put_chunk(o, "IDAT", z.data(), cl);
    put_chunk(o, "IEND", nullptr, 0);
    std::ofstream f(path, std::ios::binary);
    f.write((const char *)o.data(), o.size());
}

}  // namespace emosaic
