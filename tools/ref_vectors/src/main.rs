//! Emits reference-side known answers for the parts of emosaic's hot path whose arithmetic lives in crates that are not
//! vendored with the reference: image 0.25.2 (`Rgba::blend`, `imageops::resize` with Nearest and Lanczos3)
//! and kiddo 4.2.0 (`nearest_one` / `nearest_n` with `Manhattan` over `FixedU32<U0>`, bucket 640 — the tie-break above 320
//! tiles).  Call sites in the reference: src/main.rs:447-466 (tint), :595 and src/mosaic/tiles/utils.rs:188-189 (Lanczos3),
//! src/mosaic/tiles/tileset.rs:178-190 + src/mosaic/rendering.rs:187-195 (nearest_one), rendering.rs:307-321 (nearest_n).
//!
//!   python tools/ref_vectors/inputs.py tools/ref_vectors/inputs
//!   cargo run --release --manifest-path tools/ref_vectors/Cargo.toml -- tools/ref_vectors/inputs tests/golden/ref
//!
//! Output: one .npy per vector under tests/golden/ref/; tests/test_ref_vectors.py compares the CPU oracle and the CUDA path
//! with them when they exist.
use std::fs;
use std::io::Write;
use std::path::{Path, PathBuf};

use fixed::types::extra::U0;
use fixed::FixedU32;
use image::imageops::{self, FilterType};
use image::{Pixel, Rgb, RgbImage, Rgba, RgbaImage};
use kiddo::fixed::distance::Manhattan;
use kiddo::fixed::kdtree::KdTree;

type Fx = FixedU32<U0>;

fn write_npy(path: &Path, descr: &str, shape: &[usize], bytes: &[u8]) {
    let dims = shape.iter().map(|d| d.to_string()).collect::<Vec<_>>().join(", ");
    let tuple = if shape.len() == 1 { format!("({},)", dims) } else { format!("({})", dims) };
    let mut header = format!("{{'descr': '{}', 'fortran_order': False, 'shape': {}, }}", descr, tuple);
    while (10 + header.len() + 1) % 64 != 0 {
        header.push(' ');
    }
    header.push('\n');
    let mut f = fs::File::create(path).expect("create npy");
    f.write_all(b"\x93NUMPY\x01\x00").unwrap();
    f.write_all(&(header.len() as u16).to_le_bytes()).unwrap();
    f.write_all(header.as_bytes()).unwrap();
    f.write_all(bytes).unwrap();
}

fn le_i32(v: &[i32]) -> Vec<u8> {
    v.iter().flat_map(|x| x.to_le_bytes()).collect()
}
fn le_u32(v: &[u32]) -> Vec<u8> {
    v.iter().flat_map(|x| x.to_le_bytes()).collect()
}

/// main.rs:447-466: opaque mosaic pixel under the overlay pixel (source colour, alpha A), via imageops::overlay.
fn blend_tables(out: &Path) {
    let alphas: [u8; 8] = [1, 64, 127, 128, 200, 254, 0, 255];
    let mut table = Vec::with_capacity(alphas.len() * 65536 * 2);
    for &a in alphas.iter() {
        let mut bottom = RgbaImage::new(256, 256);
        let mut top = RgbaImage::new(256, 256);
        for bg in 0..256u32 {
            for fg in 0..256u32 {
                bottom.put_pixel(fg, bg, Rgba([bg as u8, bg as u8, bg as u8, 255]));
                top.put_pixel(fg, bg, Rgba([fg as u8, fg as u8, fg as u8, a]));
            }
        }
        imageops::overlay(&mut bottom, &top, 0, 0);
        for bg in 0..256u32 {
            for fg in 0..256u32 {
                let p = bottom.get_pixel(fg, bg);
                table.push(p[0]);
                table.push(p[3]);
            }
        }
        // the same through Pixel::blend directly must agree with overlay
        let mut q = Rgba([10u8, 10, 10, 255]);
        q.blend(&Rgba([20u8, 20, 20, a]));
        assert_eq!(q[0], bottom.get_pixel(20, 10)[0]);
    }
    write_npy(&out.join("ref_blend_alphas.npy"), "|u1", &[alphas.len()], &alphas);
    write_npy(&out.join("ref_blend.npy"), "|u1", &[alphas.len(), 256, 256, 2], &table); // [A][bg][fg] -> (value, alpha byte)
}

/// main.rs:456-461: which source column / row imageops::resize(.., Nearest) samples for every output column / row.
fn nearest_maps(out: &Path) {
    let cases: [(u32, u32); 8] = [(100, 1600), (100, 800), (1024, 32768), (37, 296), (4097, 32768), (50, 75), (640, 481), (7, 1000)];
    let mut flat: Vec<u32> = Vec::new();
    let mut index: Vec<u32> = Vec::new();
    for &(n_in, n_out) in cases.iter() {
        let img = RgbaImage::from_fn(n_in, 1, |x, _| Rgba([(x & 255) as u8, (x >> 8) as u8, 0, 255]));
        let r = imageops::resize(&img, n_out, 1, FilterType::Nearest);
        index.extend_from_slice(&[n_in, n_out, flat.len() as u32]);
        for x in 0..n_out {
            let p = r.get_pixel(x, 0);
            flat.push(p[0] as u32 | (p[1] as u32) << 8);
        }
    }
    write_npy(&out.join("ref_resize_nearest_cases.npy"), "<u4", &[cases.len(), 3], &le_u32(&index)); // (n_in, n_out, offset)
    write_npy(&out.join("ref_resize_nearest.npy"), "<u4", &[flat.len()], &le_u32(&flat));
}

/// main.rs:595 / tiles/utils.rs:188-189: imageops::resize(view, nw, nh, Lanczos3).
fn lanczos(inputs: &Path, out: &Path, name: &str, w: u32, h: u32, x0: u32, y0: u32, cw: u32, ch: u32, nw: u32, nh: u32) {
    let raw = fs::read(inputs.join(format!("lanczos_{}.bin", name))).expect("lanczos input");
    let img = RgbImage::from_raw(w, h, raw).expect("size");
    let view = imageops::crop_imm(&img, x0, y0, cw, ch);
    let r = imageops::resize(&*view, nw, nh, FilterType::Lanczos3);
    write_npy(&out.join(format!("ref_lanczos3_{}.npy", name)), "|u1", &[nh as usize, nw as usize, 3], r.as_raw());
}

/// tileset.rs:178-190 + rendering.rs:187-195 / :307-321.
fn kiddo_case<const K: usize>(inputs: &Path, out: &Path, name: &str, t: usize, q: usize) {
    let colors = fs::read(inputs.join(format!("kiddo_colors_{}.bin", name))).expect("colors");
    let queries = fs::read(inputs.join(format!("kiddo_queries_{}.bin", name))).expect("queries");
    assert_eq!(colors.len(), t * K);
    assert_eq!(queries.len(), q * K);
    let cells = K / 3;
    let cols = (cells as f64).sqrt() as usize;
    let mut tree: KdTree<Fx, i16, K, 640, u16> = KdTree::new();
    for i in 0..t {
        let mut c = [Fx::from_num(0u8); K];
        for d in 0..K {
            c[d] = Fx::from_num(colors[i * K + d]);
        }
        let idx = (i + 1) as i16;
        tree.add(&c, idx);
        // flipped_coords (tiles/utils.rs:18-43): reverse the cells inside every cell row
        let mut m = c;
        for row in 0..cols {
            for j in 0..cols / 2 {
                let a = (row * cols + j) * 3;
                let b = (row * cols + (cols - 1 - j)) * 3;
                for h in 0..3 {
                    m.swap(a + h, b + h);
                }
            }
        }
        tree.add(&m, -idx);
    }
    let mut items: Vec<i32> = Vec::with_capacity(q);
    let mut dists: Vec<u32> = Vec::with_capacity(q);
    let kn = 64usize;
    let qn = q.min(256);
    let mut n_items: Vec<i32> = vec![0; qn * kn];
    let mut n_dists: Vec<u32> = vec![u32::MAX; qn * kn];
    for i in 0..q {
        let mut c = [Fx::from_num(0u8); K];
        for d in 0..K {
            c[d] = Fx::from_num(queries[i * K + d]);
        }
        let nn = tree.nearest_one::<Manhattan>(&c);
        items.push(nn.item as i32);
        dists.push(nn.distance.to_num::<u32>());
        if i < qn {
            let list = tree.nearest_n::<Manhattan>(&c, kn);
            for (j, e) in list.iter().enumerate().take(kn) {
                n_items[i * kn + j] = e.item as i32;
                n_dists[i * kn + j] = e.distance.to_num::<u32>();
            }
        }
    }
    write_npy(&out.join(format!("ref_kiddo_{}_item.npy", name)), "<i4", &[q], &le_i32(&items));
    write_npy(&out.join(format!("ref_kiddo_{}_dist.npy", name)), "<u4", &[q], &le_u32(&dists));
    write_npy(&out.join(format!("ref_kiddo_{}_nearest_n_item.npy", name)), "<i4", &[qn, kn], &le_i32(&n_items));
    write_npy(&out.join(format!("ref_kiddo_{}_nearest_n_dist.npy", name)), "<u4", &[qn, kn], &le_u32(&n_dists));
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    if args.len() != 3 {
        eprintln!("usage: emosaic-ref-vectors <inputs dir> <output dir>");
        std::process::exit(2);
    }
    let inputs = PathBuf::from(&args[1]);
    let out = PathBuf::from(&args[2]);
    fs::create_dir_all(&out).unwrap();
    blend_tables(&out);
    nearest_maps(&out);
    let manifest = fs::read_to_string(inputs.join("manifest.txt")).expect("manifest.txt (python tools/ref_vectors/inputs.py <dir>)");
    for line in manifest.lines() {
        let f: Vec<&str> = line.split_whitespace().collect();
        if f.is_empty() {
            continue;
        }
        let n = |i: usize| f[i].parse::<u32>().unwrap();
        match f[0] {
            "lanczos" => lanczos(&inputs, &out, f[1], n(2), n(3), n(4), n(5), n(6), n(7), n(8), n(9)),
            "kiddo" => {
                let (t, q, k) = (n(2) as usize, n(3) as usize, n(4) as usize);
                match k {
                    3 => kiddo_case::<3>(&inputs, &out, f[1], t, q),
                    12 => kiddo_case::<12>(&inputs, &out, f[1], t, q),
                    27 => kiddo_case::<27>(&inputs, &out, f[1], t, q),
                    _ => panic!("unsupported vector length {}", k),
                }
            }
            _ => panic!("unknown manifest line: {}", line),
        }
    }
    let _ = Rgb([0u8, 0, 0]);
    eprintln!("wrote reference vectors to {}", out.display());
}
