// cli.cpp — `emosaic` with the reference's command line (src/main.rs:28-138) for the accelerated path, C++ host.
//   emosaic [-s N] [-o PATH] [--crop] IMG mosaic TILES_DIR [-m 1|2|...|128|1to1|4to1|random] [-f] [-t X] [--downsample K]
//           [--extensions e ...] [--gpus N]
//   emosaic [-s N] [-o PATH] [--crop] IMG prepare          (main.rs:380-386: one prepared tile)
// Tiles and the source are decoded by the minimal PNG/PPM reader of emosaic.cpp (no libjpeg in this image; the Python
// front end `python -m emosaic_b200` decodes JPEG with PIL).  A tile that is already tile_size x tile_size is taken as
// prepared (the role of the reference's ~/.cache/mosaic hit, utils.rs:73-85); any other tile goes through prepare_tile
// (trim view, crop, Lanczos3 resize on the GPU).  The source is resized like main.rs:567-595.  Cache files are
// byte-compatible.
#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>

#include "emosaic.hpp"

using namespace emosaic;

static void find_images(const std::string &dir, const std::vector<std::string> &exts, std::vector<std::string> &out) {  // image.rs:7-23
    std::vector<std::string> names;
    if (DIR *d = opendir(dir.c_str())) {
        while (dirent *e = readdir(d))
            if (strcmp(e->d_name, ".") && strcmp(e->d_name, "..")) names.push_back(e->d_name);
        closedir(d);
    }
    std::sort(names.begin(), names.end());
    for (const std::string &n : names) {
        const std::string p = dir + "/" + n;
        struct stat st;
        if (stat(p.c_str(), &st)) continue;
        if (S_ISDIR(st.st_mode)) find_images(p, exts, out);
        else {
            const size_t dot = n.find_last_of('.');
            if (dot != std::string::npos && std::find(exts.begin(), exts.end(), n.substr(dot + 1)) != exts.end()) out.push_back(p);
        }
    }
}

int main(int argc, char **argv) {
    uint32_t tile_size = 16;
    std::string output = "./output.jpg", img_path, tiles_dir, mode = "1";
    bool crop = false, force = false, mosaic = false, prepare = false, no_repeat = false;
    double tint = 0.0;
    uint32_t downsample = 1, gpus = 1;
    std::vector<std::string> exts = {"jpg", "jpeg"}, pos;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) { fprintf(stderr, "error: %s needs a value\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "-s" || a == "--tile-size") tile_size = (uint32_t)std::stoul(next());
        else if (a == "-o" || a == "--output-path") output = next();
        else if (a == "--crop") crop = true;
        else if (a == "-m" || a == "--mode") mode = next();
        else if (a == "-f" || a == "--force") force = true;
        else if (a == "-t" || a == "--tint-opacity") tint = std::stod(next());
        else if (a == "--extensions") { exts.clear(); while (i + 1 < argc && argv[i + 1][0] != '-') exts.push_back(argv[++i]); }
        else if (a == "--downsample") downsample = (uint32_t)std::stoul(next());
        else if (a == "--gpus") gpus = (uint32_t)std::stoul(next());  // row stripes / tile ranges over N GPUs (emo_group)
        else if (a == "--no-repeat") no_repeat = true;  // main.rs:663-664: render_nto1_no_repeat
        else if (a == "--randomize" || a == "--greedy" || a == "--html" || a == "--web") {
            fprintf(stderr, "error: %s is outside the accelerated path\n", a.c_str());
            return 2;
        } else if (a == "mosaic") mosaic = true;
        else if (a == "prepare") prepare = true;
        else pos.push_back(a);
    }
    if (prepare && !mosaic && pos.size() == 1) {  // main.rs:380-386: prepare_tile(&img, tile_size, crop) saved to the output path
        try {
            Context ctx(0);
            write_png(output, prepare_tile(ctx, read_image(pos[0]), tile_size, crop));
        } catch (const Error &e) {
            fprintf(stderr, "error: Failed to prepare tile from %s: %s\n", pos[0].c_str(), e.what());
            return 1;
        }
        return 0;
    }
    if (!mosaic || pos.size() != 2) {
        fprintf(stderr, "usage: emosaic [-s N] [-o PATH] [--crop] IMG mosaic TILES_DIR [-m MODE] [-f] [-t X] | IMG prepare\n");
        return 2;
    }
    if (!(tint >= 0.0 && tint <= 1.0)) { fprintf(stderr, "error: Value must be between 0 and 1\n"); return 2; }
    if (downsample < 1) { fprintf(stderr, "error: --downsample must be at least 1\n"); return 2; }
    if (gpus < 1 || gpus > 64) { fprintf(stderr, "error: --gpus must be between 1 and 64\n"); return 2; }
    img_path = pos[0];
    tiles_dir = pos[1];
    try {
        std::vector<int> devices(gpus);
        for (uint32_t i = 0; i < gpus; i++) devices[i] = (int)i;
        Context ctx(devices);
        const Image original = read_image(img_path);
        // tiles/utils.rs:63-196 from the decoded image on; `crop_tile` as prepare_tile's crop argument
        auto load_tile = [&](const std::string &p, bool crop_tile) {
            Image t = read_image(p);
            return (t.width == tile_size && t.height == tile_size) ? t : prepare_tile(ctx, t, tile_size, crop_tile);
        };
        uint32_t dim = mode == "1to1" ? 1 : mode == "4to1" ? 2 : mode == "random" ? 0 : (uint32_t)std::stoul(mode);
        if (dim == 0) {  // random mode
            std::vector<std::string> paths;
            find_images(tiles_dir, exts, paths);
            TileSet ts(1);
            for (auto &p : paths) ts.push_tile_with_image(p, {0, 0, 0}, load_tile(p, true));
            fprintf(stderr, "Tile set with %zu tiles\n", ts.len());
            write_png(output, render_random(ctx, original, ts, tile_size, 0));
            return 0;
        }
        const uint32_t N = dim * dim;
        const Image source = resize_source(ctx, original, downsample, dim);  // main.rs:567-595
        const std::string cache_path = tiles_dir + "/" + cache_file_name(N, crop);
        TileSet ts(N);
        bool have = false;
        if (!force) {
            std::ifstream f(cache_path, std::ios::binary);
            if (f) {
                std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
                try {
                    TileSet cached = deserialize_tile_set(bytes, N, &exts, true);
                    fprintf(stderr, "Reusing analysis cache\n");
                    // tileset.rs:152-155: a cache-loaded TileSet prepares its images with crop = true
                    for (const Tile &t : cached.tiles()) ts.push_tile_with_image(cached.get_path(t), t.colors, load_tile(cached.get_path(t), true));
                    have = true;
                } catch (const Error &) {
                }
            }
        }
        if (!have) {  // generate_tile_set (main.rs:740-813): one batched analysis on the GPU
            std::vector<std::string> found, paths;
            find_images(tiles_dir, exts, found);
            std::vector<Image> tiles;
            // main.rs:760-806: a tile that cannot be read or prepared is listed and left out; idx numbering follows the survivors
            std::vector<std::pair<std::string, std::string>> failed;
            for (auto &p : found) {
                try {
                    tiles.push_back(load_tile(p, crop));
                    paths.push_back(p);
                } catch (const std::exception &e) {
                    failed.emplace_back(p.compare(0, tiles_dir.size() + 1, tiles_dir + "/") == 0 ? p.substr(tiles_dir.size() + 1) : p, e.what());
                }
            }
            if (!failed.empty()) {
                fprintf(stderr, "Failed to read the following images(%zu):\n", failed.size());
                for (auto &f : failed) fprintf(stderr, "- %s: %s\n", f.first.c_str(), f.second.c_str());
            }
            const std::vector<uint8_t> colors = analyse_tiles(ctx, tiles, N);
            for (size_t i = 0; i < paths.size(); i++)  // rendering re-prepares with crop = true (tileset.rs:152-155)
                ts.push_tile_with_image(paths[i], std::vector<uint8_t>(colors.begin() + i * N * 3, colors.begin() + (i + 1) * N * 3),
                                        crop ? tiles[i] : load_tile(paths[i], true));
            const std::vector<uint8_t> blob = serialize_tile_set(ts);
            std::ofstream(cache_path, std::ios::binary).write((const char *)blob.data(), blob.size());
        }
        fprintf(stderr, "Tile set with %zu tiles\n", ts.len());
        if (no_repeat && tint > 0.0) throw Error(EMO_ERR_UNSUPPORTED, "--no-repeat with --tint-opacity is not wired in this front end");
        RenderResult r = no_repeat ? render_nto1_no_repeat(ctx, source, ts, tile_size)
                                   : render_nto1(ctx, source, ts, tile_size, false, std::nullopt, 0.0);
        if (tint > 0.0) {  // main.rs:447-478: the image as opened (not the resized copy) tints the mosaic; RGBA PNG, early return
            Image rgba(r.image.width, r.image.height, 4);
            check(emo_compose_overlay(ctx.handle(), r.item.data(), source.width, source.height, original.data.data(), original.width,
                                      original.height, tint_alpha(tint), rgba.data.data()));
            write_png(output, rgba);
            return 0;
        }
        summarise(r, ts, true, &ctx);
        write_png(output, r.image);
        const size_t dot = output.find_last_of('.');
        // the no-repeat renderer keys its statistics by output coordinates (rendering.rs:352-365): one pixel per block
        write_png((dot == std::string::npos ? output : output.substr(0, dot)) + ".stats.png",
                  render_stats(r, no_repeat ? tile_size : dim, tile_size));
    } catch (const std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
