// emosaic.hpp — C++17 host-side mirror of the reference's interface for the accelerated path, on top of
// the C ABI (include/emosaic_cuda.h).  The reference is a Rust crate (no Rust toolchain in this image), so the
// host layer that a Rust `extern "C"` binding would provide is written in C++ with the reference's names,
// argument meaning and error behaviour:
//   Tile / Tile::coords           src/mosaic/tiles/tile.rs:9-119
//   flipped_coords                src/mosaic/tiles/utils.rs:18-43
//   TileSet (push_tile, push_tile_with_image, get_tile, get_image, build_kiddo)   tiles/tileset.rs
//   analyse / get_img_colors      src/mosaic/analysis.rs:5-36
//   render_nto1 / render_random   src/mosaic/rendering.rs:124-230, :418-440
//   tint                          src/main.rs:447-478
//   adjust_dims, cache file name  src/main.rs:567-601
//   resize (Lanczos3), prepare_view / prepare_tile / rotate   src/main.rs:595, tiles/utils.rs:63-264
//   TileSet (de)serialisation     tiles/tileset.rs:28-75, tiles/tile.rs:38-65 (bincode 1.3.3 defaults)
//   RenderStats summarise/render  src/mosaic/stats.rs:87-195
// Where the reference panics or exits, these throw emosaic::Error carrying the status code and message.
// No arithmetic of the path happens here: analysis, matching and compositing are the CUDA kernels.
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/emosaic_cuda.h"

namespace emosaic {

class Error : public std::runtime_error {
public:
    Error(int code, const std::string &msg) : std::runtime_error(msg), code_(code) {}
    int code() const { return code_; }

private:
    int code_;
};

// row-major interleaved u8 image, 3 (RGB) or 4 (RGBA) channels — image::RgbImage / RgbaImage
struct Image {
    uint32_t width = 0, height = 0, channels = 3;
    std::vector<uint8_t> data;
    Image() = default;
    Image(uint32_t w, uint32_t h, uint32_t c = 3) : width(w), height(h), channels(c), data((size_t)w * h * c, 0) {}
    uint8_t *pixel(uint32_t x, uint32_t y) { return data.data() + ((size_t)y * width + x) * channels; }
    const uint8_t *pixel(uint32_t x, uint32_t y) const { return data.data() + ((size_t)y * width + x) * channels; }
    bool operator==(const Image &o) const { return width == o.width && height == o.height && channels == o.channels && data == o.data; }
};

// One GPU (emo_ctx), or several driven from this process (emo_group: library replicated by one NCCL broadcast, block rows
// of the source / tiles of the library split into contiguous ranges — the reference's rayon tasks, rendering.rs:68-89 and
// main.rs:760-794 — every GPU copying its stripe of the result straight into the caller's image).
class Context {
public:
    explicit Context(int device = 0);
    explicit Context(const std::vector<int> &devices);  // devices.size() > 1: a group; analyse_tiles / render_nto1 shard over it
    ~Context();
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    emo_ctx *handle() const { return h_; }      // the (first) GPU: every single-GPU call runs here
    emo_group *group() const { return g_; }     // nullptr for a single GPU
    int gpus() const { return g_ ? emo_group_size(g_) : 1; }
    // 1to1 search index (include/emosaic_cuda.h §2b): built on demand by the match; these force or forbid it
    void build_index() const;
    void set_match_mode(int mode) const;  // EMO_MATCH_AUTO | EMO_MATCH_SCAN | EMO_MATCH_INDEX

private:
    emo_ctx *h_ = nullptr;
    emo_group *g_ = nullptr;
};

void check(int rc);  // throws Error(rc, emo_last_error()) when rc != 0

// ---- tiles/utils.rs:18-43 ------------------------------------------------------------------------
void flipped_coords(std::vector<uint32_t> &coords);

// ---- tiles/tile.rs --------------------------------------------------------------------------------
struct Tile {
    std::vector<uint8_t> colors;  // [N][3]
    uint16_t idx = 0;
    bool flipped = false;
    std::optional<std::string> date_taken;
    static Tile from_colors(std::vector<uint8_t> colors) { return Tile{std::move(colors), 0, false, std::nullopt}; }
    std::vector<uint32_t> coords() const;  // tile.rs:106-119
};

// ---- analysis.rs ----------------------------------------------------------------------------------
std::vector<uint8_t> analyse(Context &ctx, const Image &img, uint32_t N);                        // analysis.rs:5-20
std::vector<uint8_t> analyse_tiles(Context &ctx, const std::vector<Image> &tiles, uint32_t N);  // main.rs:786-794, one batch
std::vector<uint8_t> get_img_colors(uint32_t x, uint32_t y, uint32_t step, const Image &src, uint32_t N);  // analysis.rs:23-36

// ---- tiles/tileset.rs -----------------------------------------------------------------------------
class TileSet {
public:
    explicit TileSet(uint32_t N) : N_(N) {}
    uint32_t cells() const { return N_; }
    size_t len() const { return tiles_.size(); }
    void push_tile(const std::string &path, std::vector<uint8_t> colors, std::optional<std::string> date = std::nullopt);
    void push_tile_with_image(const std::string &path, std::vector<uint8_t> colors, Image image);
    std::optional<Tile> get_tile(int32_t idx) const;                 // tileset.rs:131-143
    Image get_image(const Tile &tile, uint32_t tile_size) const;      // tileset.rs:146-161 (in-memory images only)
    const std::string &get_path(const Tile &tile) const { return paths_.at(tile.idx - 1); }
    void build_kiddo(Context &ctx, uint32_t tile_size) const;         // tileset.rs:178-190 -> emo_set_library
    const std::vector<Tile> &tiles() const { return tiles_; }
    const std::vector<std::string> &paths() const { return paths_; }

private:
    uint32_t N_;
    std::vector<Tile> tiles_;
    std::vector<std::string> paths_;
    std::vector<Image> images_;  // index idx-1; empty Image when absent
};

// ---- rendering.rs ---------------------------------------------------------------------------------
struct RenderResult {
    Image image;
    std::vector<int32_t> item;   // [bh][bw] what RenderStats::push_tile receives (tile id, signed)
    std::vector<uint32_t> dist;  // and its distance
    uint32_t bw = 0, bh = 0;
};
RenderResult render_nto1(Context &ctx, const Image &source, const TileSet &tile_set, uint32_t tile_size, bool no_repeat = false,
                         std::optional<double> randomize = std::nullopt, double tint_opacity = 0.0);
// rendering.rs:262-401 (emo_no_repeat); `page` = candidates per block of the first page, 0 = automatic (lists are refilled on demand)
RenderResult render_nto1_no_repeat(Context &ctx, const Image &source, const TileSet &tile_set, uint32_t tile_size, uint32_t page = 0);
Image render_random(Context &ctx, const Image &source, const TileSet &tile_set, uint32_t tile_size, uint64_t seed);
uint8_t tint_alpha(double tint_opacity);  // main.rs:449

// ---- main.rs:567-601 --------------------------------------------------------------------------------
std::pair<uint32_t, uint32_t> adjust_dims(uint32_t w, uint32_t h, uint32_t downsample, uint32_t dim);

// ---- image 0.25.2 imageops::resize(.., Lanczos3) and prepare_tile (tiles/utils.rs:63-264) ---------------
struct View { uint32_t x = 0, y = 0, w = 0, h = 0; };
Image resize_lanczos3(Context &ctx, const Image &img, uint32_t nw, uint32_t nh, std::optional<View> view = std::nullopt);  // emo_resize
Image resize_source(Context &ctx, const Image &original, uint32_t downsample, uint32_t dim);  // main.rs:567-595
uint32_t most_common_value(const std::vector<uint32_t> &values);                              // utils.rs:262-273 (ties: smallest)
View prepare_view(const Image &img, uint32_t tile_size, bool crop);                           // utils.rs:93-186
Image rotate(const Image &img, uint32_t orientation);                                         // utils.rs:248-264
Image prepare_tile(Context &ctx, const Image &decoded, uint32_t tile_size, bool crop, uint32_t orientation = 1);  // utils.rs:63-196
std::string cache_file_name(uint32_t N, bool crop);

// ---- cache (bincode 1.3.3 defaults) -----------------------------------------------------------------
std::vector<uint8_t> serialize_tile_set(const TileSet &ts);
// main.rs:617-654: drops entries whose extension is not allowed / whose file is missing, renumbers 1..n
TileSet deserialize_tile_set(const std::vector<uint8_t> &bytes, uint32_t N, const std::vector<std::string> *extensions = nullptr,
                             bool check_exists = false);

// ---- stats.rs ---------------------------------------------------------------------------------------
struct StatsSummary {
    size_t total = 0, unique = 0;
    double average_distance = 0;
    std::vector<std::pair<std::string, uint32_t>> top, worst;
};
// ctx: reduce the counts and sums on the GPU (emo_stats) instead of on the host; the two top-10 lists are host work either way
StatsSummary summarise(const RenderResult &r, const TileSet &ts, bool print = true, Context *ctx = nullptr);
Image render_stats(const RenderResult &r, uint32_t dim, uint32_t tile_size);

// ---- minimal image I/O for the CLI (PPM P6 and non-interlaced 8-bit PNG over zlib) ---------------------
Image read_image(const std::string &path);  // -> RGB
void write_png(const std::string &path, const Image &img);
void write_ppm(const std::string &path, const Image &img);

}  // namespace emosaic
