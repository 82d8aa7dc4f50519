// host_tests.cpp — the reference's own unit tests for the hot path, restated against the C++ host mirror
// (emosaic.hpp) so that they read like the originals.  `host_tests --cpu` runs only the tests that need no GPU.
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>

#include "emosaic.hpp"

using namespace emosaic;

static int failures = 0;
#define CHECK(cond)                                                         \
    do {                                                                    \
        if (!(cond)) {                                                      \
            fprintf(stderr, "  CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
            failures++;                                                     \
        }                                                                   \
    } while (0)

static void run(const char *name, const std::function<void()> &f) {
    const int before = failures;
    try {
        f();
    } catch (const std::exception &e) {
        fprintf(stderr, "  exception: %s\n", e.what());
        failures++;
    }
    printf("%s %s\n", failures == before ? "PASS" : "FAIL", name);
}

template <typename F>
static bool throws_with(F f, const char *needle) {
    try {
        f();
    } catch (const Error &e) {
        return strstr(e.what(), needle) != nullptr;
    }
    return false;
}

static std::vector<Image> universe(uint32_t N) {  // mod.rs:83-106: all black&white dim x dim images but the all-white one
    uint32_t dim = 1;
    while (dim * dim < N) dim++;
    std::vector<Image> u;
    for (uint32_t index = 0; index < (1u << N) - 1; index++) {
        Image img(dim, dim, 3);
        for (uint32_t y = 0; y < dim; y++)
            for (uint32_t x = 0; x < dim; x++) {
                const uint32_t i = y * dim + x;              // bits are reversed in the reference: bit (N-1-i)
                const bool white = (index & (1u << (N - 1 - i))) != 0;
                memset(img.pixel(x, y), white ? 255 : 0, 3);
            }
        u.push_back(img);
    }
    return u;
}

int main(int argc, char **argv) {
    const bool cpu_only = argc > 1 && !strcmp(argv[1], "--cpu");

    // tiles/tile.rs:127-140
    run("test_tile_coords", [] {
        Tile t = Tile::from_colors({1, 2, 3});
        CHECK((t.coords() == std::vector<uint32_t>{1, 2, 3}));
        Tile t4 = Tile::from_colors({1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12});
        CHECK((t4.coords() == std::vector<uint32_t>{1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12}));
        t4.flipped = true;
        CHECK((t4.coords() == std::vector<uint32_t>{4, 5, 6, 1, 2, 3, 10, 11, 12, 7, 8, 9}));
    });
    // tiles/utils.rs:302-308
    run("test_flipped_coords", [] {
        std::vector<uint32_t> c = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12};
        flipped_coords(c);
        CHECK((c == std::vector<uint32_t>{4, 5, 6, 1, 2, 3, 10, 11, 12, 7, 8, 9}));
        flipped_coords(c);
        CHECK((c == std::vector<uint32_t>{1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12}));
    });
    // analysis.rs:58-71
    run("test_get_img_colors", [] {
        Image img(4, 4, 3);
        for (uint32_t y = 0; y < 4; y++)
            for (uint32_t x = 0; x < 4; x++) {
                uint8_t *p = img.pixel(x, y);
                p[0] = x * 64; p[1] = y * 64; p[2] = 128;
            }
        CHECK((get_img_colors(0, 0, 2, img, 4) == std::vector<uint8_t>{0, 0, 128, 64, 0, 128, 0, 64, 128, 64, 64, 128}));
    });
    // mod.rs:26-46
    run("test_tile_set_new_push", [] {
        TileSet ts(1);
        CHECK(ts.len() == 0);
        ts.push_tile("", {0, 0, 0});
        CHECK(ts.len() == 1);
        CHECK(ts.get_tile(1)->idx == 1 && !ts.get_tile(1)->flipped && ts.get_tile(-1)->flipped);
        CHECK(!ts.get_tile(0) && !ts.get_tile(2));
    });
    // main.rs:567-601
    run("test_adjust_dims_and_cache_name", [] {
        CHECK((adjust_dims(101, 103, 1, 2) == std::pair<uint32_t, uint32_t>{100, 102}));
        CHECK((adjust_dims(1023, 77, 2, 3) == std::pair<uint32_t, uint32_t>{510, 39}));
        CHECK(cache_file_name(1, false) == ".emosaic_1to1" && cache_file_name(4, true) == ".emosaic_4to1_cropped");
        CHECK(tint_alpha(0.5) == 127 && tint_alpha(1.0) == 255 && tint_alpha(0.0) == 0);
    });
    // tileset.rs:28-75 + main.rs:624-654
    run("test_cache_roundtrip", [] {
        TileSet ts(4);
        for (int t = 0; t < 9; t++) {
            std::vector<uint8_t> c(12);
            for (int i = 0; i < 12; i++) c[i] = (uint8_t)(t * 12 + i);
            ts.push_tile("/tmp/x/img" + std::to_string(t) + (t % 3 ? ".jpg" : ".png"), c,
                         t % 2 ? std::optional<std::string>("2021:01:0" + std::to_string(t)) : std::nullopt);
        }
        const std::vector<uint8_t> blob = serialize_tile_set(ts);
        CHECK(blob.size() > 16 && blob[0] == 9 && blob[8] == 12 && blob[16] == 0 && blob[28] == 1 && blob[29] == 0 && blob[30] == 0);
        TileSet back = deserialize_tile_set(blob, 4);
        CHECK(back.len() == 9 && back.paths() == ts.paths() && back.tiles()[3].colors == ts.tiles()[3].colors);
        CHECK(back.tiles()[1].date_taken == ts.tiles()[1].date_taken && !back.tiles()[0].date_taken);
        std::vector<std::string> exts = {"jpg"};
        TileSet f = deserialize_tile_set(blob, 4, &exts);
        CHECK(f.len() == 6 && f.tiles()[0].idx == 1 && f.tiles()[0].colors == ts.tiles()[1].colors);  // renumbered 1..n
        CHECK(throws_with([&] { deserialize_tile_set(blob, 1); }, "wrong vector length"));
        std::vector<uint8_t> cut(blob.begin(), blob.end() - 3);
        CHECK(throws_with([&] { deserialize_tile_set(cut, 4); }, "truncated"));
    });
    // stats.rs:87-195
    run("test_stats", [] {
        RenderResult r;
        r.bw = 3; r.bh = 2;
        r.item = {1, -2, 1, 3, 1, 2};
        r.dist = {5, 9, 0, 7, 7, 1};
        TileSet ts(1);
        for (const char *p : {"a", "b", "c"}) ts.push_tile(p, {0, 0, 0});
        StatsSummary s = summarise(r, ts, false);
        CHECK(s.total == 6 && s.unique == 3 && std::abs(s.average_distance - 29.0 / 6) < 1e-12);
        CHECK(s.top[0].first == "a" && s.top[0].second == 3 && s.worst[0].first == "b" && s.worst[0].second == 9);
        Image im = render_stats(r, 1, 1);
        CHECK(im.width == 3 && im.height == 2 && im.pixel(0, 0)[0] == 141 && im.pixel(1, 0)[0] == 255 && im.pixel(2, 1)[0] == 28);
    });
    // the reference's own stats tests, stats.rs:247-317 (empty summary, the two panics, test_render_basic), and unplaced blocks
    run("test_stats_reference_cases", [] {
        TileSet ts(1);
        for (const char *p : {"test1.jpg", "test2.jpg"}) ts.push_tile(p, {255, 0, 0});
        RenderResult empty;
        CHECK(summarise(empty, ts, false).total == 0);                                       // test_summarise_empty: no panic
        CHECK(throws_with([&] { render_stats(empty, 1, 16); }, "Cannot render visualization: no tiles recorded"));
        RenderResult one;
        one.bw = one.bh = 1; one.item = {1}; one.dist = {100};
        CHECK(throws_with([&] { render_stats(one, 1, 0); }, "Tile size must be greater than 0"));
        RenderResult r;  // test_summarise_with_tiles: tiles at (0,0) d=10, (10,10) d=20, (20,20) d=15; tile 1 used twice
        r.bw = 3; r.bh = 1; r.item = {1, 2, 1}; r.dist = {10, 20, 15};
        StatsSummary s = summarise(r, ts, false);
        CHECK(s.total == 3 && s.unique == 2 && s.average_distance == 15.0 && s.top[0].first == "test1.jpg" && s.top[0].second == 2);
        CHECK(s.worst[0].first == "test2.jpg" && s.worst[0].second == 20 && s.worst[2].second == 10);
        RenderResult d;  // test_render_basic: (0,0) d=50 and (16,16) d=150 at tile size 16 -> 2x2, darker where the match is better
        d.bw = d.bh = 2; d.item = {1, 0, 0, 1}; d.dist = {50, 0, 0, 150};
        Image im = render_stats(d, 16, 16);
        CHECK(im.width == 2 && im.height == 2 && im.pixel(0, 0)[0] < im.pixel(1, 1)[0] && im.pixel(1, 1)[0] == 255 && im.pixel(0, 0)[0] == 85);
        // a no-repeat render that ran out of tiles (item 0, rendering.rs:347-351): not counted, no out_of_range
        RenderResult u;
        u.bw = 3; u.bh = 1; u.item = {2, 0, -1}; u.dist = {4, 999, 6};
        StatsSummary su = summarise(u, ts, false);
        CHECK(su.total == 2 && su.unique == 2 && su.average_distance == 5.0 && su.worst[0].second == 6);
        CHECK(render_stats(u, 8, 8).width == 3 && render_stats(u, 8, 8).pixel(1, 0)[0] == 0);
    });
    // corrupt cache files fail with an Error before any out-of-bounds read or huge allocation (bincode Err -> re-analysis)
    run("test_cache_corrupt_lengths", [] {
        TileSet ts(1);
        ts.push_tile("/tmp/a.jpg", {1, 2, 3}, std::optional<std::string>("2020:01:01"));
        std::vector<uint8_t> blob = serialize_tile_set(ts);
        std::vector<uint8_t> huge_T = blob;
        for (int i = 0; i < 8; i++) huge_T[i] = 0xFF;
        CHECK(throws_with([&] { deserialize_tile_set(huge_T, 1); }, "truncated"));
        std::vector<uint8_t> huge_date = blob;  // date length field: 8 (T) + 8 (len) + 3 + 2 (idx) + 1 (tag) = offset 22
        for (int i = 0; i < 8; i++) huge_date[22 + i] = 0xFF;
        CHECK(throws_with([&] { deserialize_tile_set(huge_date, 1); }, "truncated"));
        std::vector<uint8_t> huge_path = blob;
        for (int i = 0; i < 8; i++) huge_path[blob.size() - 10 - 8 + i] = 0xFF;  // "/tmp/a.jpg" = 10 bytes, its length right before
        CHECK(throws_with([&] { deserialize_tile_set(huge_path, 1); }, "truncated"));
    });
    // PNG/PPM codec used by the CLI
    run("test_image_io_roundtrip", [] {
        Image im(5, 3, 3);
        for (size_t i = 0; i < im.data.size(); i++) im.data[i] = (uint8_t)(i * 37);
        write_png("/tmp/emosaic_host_test.png", im);
        CHECK(read_image("/tmp/emosaic_host_test.png") == im);
        write_ppm("/tmp/emosaic_host_test.ppm", im);
        CHECK(read_image("/tmp/emosaic_host_test.ppm") == im);
    });
    // tiles/utils.rs:284-289 (test_most_common_value) + the trim view / rotation bookkeeping of prepare_tile (:93-186, :248-264)
    run("test_most_common_value / prepare_view / rotate", [] {
        CHECK(most_common_value({1, 2, 2, 3, 3, 3, 4}) == 3);
        CHECK(most_common_value({}) == 0 && most_common_value({9, 4, 9, 4}) == 4);
        Image img(60, 40, 3);  // white frame: 5 left, 3 right, 4 top, 6 bottom
        for (uint32_t y = 0; y < 40; y++)
            for (uint32_t x = 0; x < 60; x++) memset(img.pixel(x, y), (x >= 5 && x < 57 && y >= 4 && y < 34) ? (int)((x * 7 + y * 3) % 200) : 255, 3);
        const View v = prepare_view(img, 8, false);  // [first, last): the last non-white column / row is left out
        CHECK(v.x == 5 && v.y == 4 && v.w == 51 && v.h == 29);
        const View c = prepare_view(img, 8, true);
        CHECK(c.w == 29 && c.h == 29 && c.x == 5 + 11 && c.y == 4);
        Image white(20, 20, 3);
        memset(white.data.data(), 255, white.data.size());
        CHECK(throws_with([&] { prepare_view(white, 8, false); }, "assertion failed: first_non_white_col < last_non_white_col"));
        CHECK(throws_with([&] { prepare_view(Image(30, 7, 3), 8, false); }, "smaller than the tile size"));
        Image a(3, 2, 3);
        for (size_t i = 0; i < a.data.size(); i++) a.data[i] = (uint8_t)i;
        CHECK(rotate(a, 1) == a && rotate(rotate(a, 6), 8) == a && rotate(rotate(a, 3), 3) == a && rotate(rotate(a, 7), 7) == a);
        const Image r6 = rotate(a, 6);  // clockwise: top-left goes to top-right
        CHECK(r6.width == 2 && r6.height == 3 && !memcmp(r6.pixel(1, 0), a.pixel(0, 0), 3) && !memcmp(r6.pixel(0, 0), a.pixel(0, 1), 3));
        const Image r5 = rotate(a, 5);  // transpose
        CHECK(!memcmp(r5.pixel(1, 2), a.pixel(2, 1), 3));
    });
    if (cpu_only) {
        printf("%s (%d failures, cpu-only subset)\n", failures ? "FAILED" : "OK", failures);
        return failures ? 1 : 0;
    }

    Context ctx(0);
    // color.rs:49-64 + analysis.rs:44-55 through the kernel
    run("test_average_color_basic / test_analyse_single_color", [&] {
        Image img(2, 2, 3);
        const uint8_t px[4][3] = {{100, 150, 200}, {200, 100, 50}, {50, 200, 100}, {150, 50, 150}};
        for (int i = 0; i < 4; i++) memcpy(img.pixel(i % 2, i / 2), px[i], 3);
        CHECK((analyse(ctx, img, 1) == std::vector<uint8_t>{125, 125, 125}));
        Image red(2, 2, 3);
        for (int i = 0; i < 4; i++) red.pixel(i % 2, i / 2)[0] = 255;
        CHECK((analyse(ctx, red, 4) == std::vector<uint8_t>{255, 0, 0, 255, 0, 0, 255, 0, 0, 255, 0, 0}));
        Image tiny(2, 2, 3);
        CHECK(throws_with([&] { analyse(ctx, tiny, 9); }, "Rectangle dimensions must be positive"));  // color.rs:18
    });
    // mod.rs:48-68
    run("test_render_random / test_render_nto1 (dimensions)", [&] {
        Image source(5, 2, 3);
        TileSet ts(1);
        ts.push_tile_with_image("", {0, 0, 0}, Image(8, 8, 3));
        RenderResult out = render_nto1(ctx, source, ts, 8);
        CHECK(out.image.width == 5 * 8 && out.image.height == 2 * 8);
        Image rnd = render_random(ctx, Image(10, 10, 3), ts, 8, 1);
        CHECK(rnd.width == 80 && rnd.height == 80);
    });
    // mod.rs:83-161
    for (uint32_t N : {1u, 4u, 9u}) {
        run(("test_analyse_tiles_consistency_" + std::to_string(N)).c_str(), [&] {
            uint32_t dim = 1;
            while (dim * dim < N) dim++;
            const std::vector<Image> uni = universe(N);
            const std::vector<uint8_t> colors = analyse_tiles(ctx, uni, N);
            TileSet ts(N);
            for (size_t i = 0; i < uni.size(); i++)
                ts.push_tile_with_image("", std::vector<uint8_t>(colors.begin() + i * N * 3, colors.begin() + (i + 1) * N * 3), uni[i]);
            size_t shown = 0;
            for (const Image &img : uni) {  // every member renders to itself
                CHECK(render_nto1(ctx, img, ts, dim).image == img);
                if (++shown >= 64) break;
            }
            for (size_t a = 0; a + 1 < uni.size() && a < 64; a += 2) {  // 1x2 stacks of consecutive members
                Image img(dim, 2 * dim, 3);
                memcpy(img.data.data(), uni[a].data.data(), uni[a].data.size());
                memcpy(img.data.data() + uni[a].data.size(), uni[a + 1].data.data(), uni[a + 1].data.size());
                RenderResult r = render_nto1(ctx, img, ts, dim);
                CHECK(r.image == img);
                CHECK(r.dist[0] == 0 && r.dist[1] == 0);
            }
            CHECK(throws_with([&] { render_nto1(ctx, uni[0], ts, dim, true); }, "outside the accelerated"));
            // mod.rs:122-126 and :140-144: the same round trips through render_nto1_no_repeat
            shown = 0;
            for (const Image &img : uni) {
                CHECK(render_nto1_no_repeat(ctx, img, ts, dim).image == img);
                if (++shown >= 32) break;
            }
            for (size_t a = 0; a + 1 < uni.size() && a < 32; a += 2) {
                Image img(dim, 2 * dim, 3);
                memcpy(img.data.data(), uni[a].data.data(), uni[a].data.size());
                memcpy(img.data.data() + uni[a].data.size(), uni[a + 1].data.data(), uni[a + 1].data.size());
                RenderResult r = render_nto1_no_repeat(ctx, img, ts, dim, 1);  // one candidate per page: every conflict refills
                CHECK(r.image == img);
                CHECK(r.item[0] != r.item[1] && r.item[0] != -r.item[1]);
            }
        });
    }
    // the 1to1 search index answers exactly like the scan (include/emosaic_cuda.h §2b)
    run("test_search_index_equals_scan", [&] {
        TileSet ts(1);
        uint32_t seed = 12345;
        auto rnd = [&] { seed = seed * 1664525u + 1013904223u; return (uint8_t)(seed >> 24); };
        for (int i = 0; i < 500; i++) {
            const uint8_t r = rnd(), g = rnd(), b = rnd();
            Image tile(8, 8, 3);
            for (size_t k = 0; k < tile.data.size(); k += 3) { tile.data[k] = r; tile.data[k + 1] = g; tile.data[k + 2] = b; }
            ts.push_tile_with_image("", {r, g, b}, tile);
        }
        Image src(64, 48, 3);
        for (auto &v : src.data) v = rnd();
        ctx.set_match_mode(EMO_MATCH_SCAN);
        RenderResult a = render_nto1(ctx, src, ts, 8);
        ctx.set_match_mode(EMO_MATCH_INDEX);
        RenderResult b = render_nto1(ctx, src, ts, 8);
        ctx.set_match_mode(EMO_MATCH_AUTO);
        CHECK(a.image == b.image && a.item == b.item && a.dist == b.dist);
        TileSet four(4);
        four.push_tile_with_image("", std::vector<uint8_t>(12, 0), Image(8, 8, 3));
        four.build_kiddo(ctx, 8);
        CHECK(throws_with([&] { ctx.build_index(); }, "N == 1"));
    });
    // rendering.rs:292-298 and :347-351
    run("test_no_repeat_rules", [&] {
        TileSet ts(1);
        for (int i = 0; i < 3; i++) {
            Image tile(4, 4, 3);
            for (auto &v : tile.data) v = (uint8_t)(i * 100);
            ts.push_tile_with_image("", {(uint8_t)(i * 100), (uint8_t)(i * 100), (uint8_t)(i * 100)}, tile);
        }
        CHECK(throws_with([&] { render_nto1_no_repeat(ctx, Image(7, 1, 3), ts, 4); }, "Insufficient tiles for no-repeat mode: need 7 tiles but only have 6"));
        RenderResult r = render_nto1_no_repeat(ctx, Image(5, 1, 3), ts, 4);  // 5 black blocks, 3 tiles: two blocks stay unplaced
        int placed = 0;
        for (int32_t it : r.item) placed += it != 0;
        CHECK(placed == 3 && r.item[0] == 1 && r.dist[0] == 0);
        // the summary with its sums reduced on the GPU (emo_stats) equals the host-only one; unplaced blocks have no entry
        TileSet named(1);
        for (const char *p : {"a", "b", "c"}) named.push_tile(p, {0, 0, 0});
        const StatsSummary host = summarise(r, named, false), dev = summarise(r, named, false, &ctx);
        CHECK(dev.total == 3 && dev.total == host.total && dev.unique == host.unique && dev.average_distance == host.average_distance);
        CHECK(dev.top == host.top && dev.worst == host.worst);
    });
    // main.rs:603-615 exits become errors
    run("test_dimension_rules", [&] {
        TileSet ts(4);
        ts.push_tile_with_image("", std::vector<uint8_t>(12, 0), Image(8, 8, 3));
        CHECK(throws_with([&] { render_nto1(ctx, Image(5, 4, 3), ts, 8); }, "Dimensions must be divisible by 2"));
        CHECK(throws_with([&] { render_nto1(ctx, Image(4, 4, 3), ts, 7); }, "Tile size must be divisible by 2"));
    });
    // tint: alpha byte and the SURVEY §8c known answers (A = 127)
    run("test_tint_known_answers", [&] {
        TileSet ts(1);
        Image tile(4, 4, 3);
        const uint8_t bg[6] = {0, 255, 100, 10, 255, 128}, fg[6] = {255, 0, 200, 20, 255, 128}, want[6] = {127, 127, 149, 14, 255, 128};
        for (int k = 0; k < 6; k++) {
            for (auto &v : tile.data) v = bg[k];
            TileSet one(1);
            one.push_tile_with_image("", {bg[k], bg[k], bg[k]}, tile);
            Image src(1, 1, 3);
            memset(src.data.data(), fg[k], 3);
            RenderResult r = render_nto1(ctx, src, one, 4, false, std::nullopt, 0.5);
            CHECK(r.image.channels == 4 && r.image.pixel(2, 1)[0] == want[k] && r.image.pixel(0, 0)[3] == 255);
        }
    });
    // image 0.25.2 imageops::resize(Lanczos3) through emo_resize (main.rs:595, utils.rs:188-189) and utils.rs:291-299
    run("test_resize / test_prepare_tile", [&] {
        Image flat(131, 97, 3);
        for (size_t k = 0; k < flat.data.size(); k += 3) { flat.data[k] = 7; flat.data[k + 1] = 200; flat.data[k + 2] = 255; }
        const Image small = resize_lanczos3(ctx, flat, 16, 12);
        bool same = small.width == 16 && small.height == 12;
        for (size_t k = 0; k < small.data.size(); k += 3) same = same && small.data[k] == 7 && small.data[k + 1] == 200 && small.data[k + 2] == 255;
        CHECK(same);  // normalised weights keep a flat image flat
        Image img(45, 37, 3);
        uint32_t seed = 99;
        for (auto &v : img.data) { seed = seed * 1664525u + 1013904223u; v = (uint8_t)((seed >> 24) % 230); }
        CHECK(resize_lanczos3(ctx, img, 45, 37) == img);  // same dimensions: copy (sample.rs resize())
        const Image src = resize_source(ctx, img, 1, 2);  // 45 x 37 -> 44 x 36
        CHECK(src.width == 44 && src.height == 36);
        const Image tile = prepare_tile(ctx, img, 32, true);  // utils.rs:291-299: prepare_tile(.., 32, true) is 32 x 32
        CHECK(tile.width == 32 && tile.height == 32);
        CHECK(prepare_tile(ctx, img, 32, true, 6) == rotate(tile, 6));
        CHECK(throws_with([&] { resize_lanczos3(ctx, img, 8, 8, View{40, 0, 10, 10}); }, "extends beyond"));
    });
    printf("%s (%d failures)\n", failures ? "FAILED" : "OK", failures);
    return failures ? 1 : 0;
}
