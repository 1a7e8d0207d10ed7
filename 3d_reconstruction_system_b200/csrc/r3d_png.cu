// r3d_png.cu -- a1: depth / disparity PNG decode on a host thread pool (host code only; nvcc compiles it with the rest).
//
// Replaces cv.imread(path, IMREAD_GRAYSCALE) (transfer/camera_to_world.py:160) and cv.imread(path, IMREAD_UNCHANGED)
// + [:, :, 1] (transfer/pixel_to_camera.py:133-134) for PNG files, for whole frame batches at once: one worker thread per
// core inflates (zlib), un-filters and converts straight into the caller's (pinned) frame stack, which is what the
// back-projection kernel uploads.  The reference decodes one frame at a time on one core (3 ms per 1242x375 frame),
// which is 2 500 x the time the GPU needs for that frame.
//
// Conversions are pinned against OpenCV 4.13 / libpng in tests/test_png_cpu.py:
//   GRAY8   : what IMREAD_GRAYSCALE returns: 16-bit samples >> 8, colour through libpng's rgb_to_gray
//             (8-bit: (9797 R + 19234 G + 3737 B) >> 15, truncating; 16-bit: (... + 16384) >> 15, then >> 8),
//             alpha dropped, 1/2/4-bit grey expanded to 0..255, palette through its RGB entries
//   CHANNEL : what IMREAD_UNCHANGED[:, :, c] returns: c indexes B, G, R, A (grey+alpha expands to B=G=R=grey);
//             a single-channel file has no channel axis and is an error, like the reference's IndexError
//   RAW     : IMREAD_UNCHANGED with the first channel taken when there are several (formats.imread_raw)
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "r3d_common.cuh"

namespace {

using r3d::set_error;

struct PngHeader {
    uint32_t W = 0, H = 0;
    int depth = 0, color = 0, interlace = 0;
    int channels() const { return color == 0 ? 1 : color == 2 ? 3 : color == 3 ? 1 : color == 4 ? 2 : 4; }
    int cv_channels() const { return (color == 0) ? 1 : (color == 2 ? 3 : (color == 3 ? 3 : 4)); }   // shape[2] of IMREAD_UNCHANGED (1 = no axis)
    int out_depth() const { return depth == 16 ? 16 : 8; }
};

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

bool read_file(const char* path, std::vector<uint8_t>& buf, std::string& err) {
    FILE* f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n < 0) { fclose(f); err = std::string("cannot size ") + path; return false; }
    buf.resize((size_t)n);
    const size_t got = n ? fread(buf.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    if (got != (size_t)n) { err = std::string("short read on ") + path; return false; }
    return true;
}

// header + concatenated IDAT + palette
bool parse(const std::vector<uint8_t>& file, PngHeader& h, std::vector<uint8_t>& idat, std::vector<uint8_t>& plte, bool& has_trns, std::string& err) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 + 25 || memcmp(file.data(), sig, 8) != 0) { err = "not a PNG file"; return false; }
    size_t p = 8;
    bool seen_ihdr = false;
    has_trns = false;
    while (p + 12 <= file.size()) {
        const uint32_t len = be32(&file[p]);
        const uint8_t* type = &file[p + 4];
        if (p + 12 + (size_t)len > file.size()) { err = "truncated chunk"; return false; }
        const uint8_t* d = &file[p + 8];
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) { err = "bad IHDR"; return false; }
            h.W = be32(d); h.H = be32(d + 4); h.depth = d[8]; h.color = d[9]; h.interlace = d[12];
            if (d[10] != 0 || d[11] != 0) { err = "unknown compression / filter method"; return false; }
            seen_ihdr = true;
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), d, d + len);
        } else if (!memcmp(type, "PLTE", 4)) {
            plte.assign(d, d + len);
        } else if (!memcmp(type, "tRNS", 4)) {
            has_trns = true;
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        p += 12 + (size_t)len;
    }
    if (!seen_ihdr) { err = "no IHDR"; return false; }
    if (h.W == 0 || h.H == 0 || h.W > 65535 || h.H > 65535) { err = "unsupported image size"; return false; }
    const int d = h.depth, c = h.color;
    const bool ok = (c == 0 && (d == 1 || d == 2 || d == 4 || d == 8 || d == 16)) || ((c == 2 || c == 4 || c == 6) && (d == 8 || d == 16)) ||
                    (c == 3 && (d == 1 || d == 2 || d == 4 || d == 8));
    if (!ok) { err = "invalid colour type / bit depth"; return false; }
    if (h.interlace) { err = "interlaced PNG is not supported"; return false; }
    if (c == 3 && plte.size() < 3) { err = "palette image without PLTE"; return false; }
    return true;
}

inline int paeth(int a, int b, int c) {
    const int p = a + b - c;
    const int pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Sub filter: cur[i] += cur[i - bpp].  Written naively the chain runs through memory (store -> load forwarding, ~5 cycles
// per byte and lane); with the bpp running values carried in registers it is one add per byte and lane.
template <int BPP>
inline void unfilter_sub_n(uint8_t* cur, size_t rb) {
    uint8_t acc[BPP];
    for (int k = 0; k < BPP; ++k) acc[k] = 0;
    size_t i = 0;
    for (; i + BPP <= rb; i += BPP)
        for (int k = 0; k < BPP; ++k) { acc[k] = (uint8_t)(acc[k] + cur[i + k]); cur[i + k] = acc[k]; }
    for (int k = 0; i + k < rb; ++k) cur[i + k] = (uint8_t)(acc[k] + cur[i + k]);
}
// Average and Paeth with the row above: the same idea, the left (and upper-left) neighbours of every lane stay in registers
template <int BPP>
inline void unfilter_avg_n(uint8_t* cur, const uint8_t* up, size_t rb) {
    unsigned a[BPP];
    for (int k = 0; k < BPP; ++k) a[k] = 0;
    size_t i = 0;
    for (; i + BPP <= rb; i += BPP)
        for (int k = 0; k < BPP; ++k) { a[k] = (cur[i + k] + ((a[k] + up[i + k]) >> 1)) & 255u; cur[i + k] = (uint8_t)a[k]; }
    for (int k = 0; i + k < rb; ++k) cur[i + k] = (uint8_t)(cur[i + k] + ((a[k] + up[i + k]) >> 1));
}
template <int BPP>
inline void unfilter_paeth_n(uint8_t* cur, const uint8_t* up, size_t rb) {
    int a[BPP], c[BPP];
    for (int k = 0; k < BPP; ++k) a[k] = c[k] = 0;
    size_t i = 0;
    auto one = [&](size_t idx, int k) {
        const int b = up[idx];
        const int pb = a[k] - c[k], pa = b - c[k];           // p - b, p - a with p = a + b - c
        const int pc = pa + pb;
        const int apa = pa < 0 ? -pa : pa, apb = pb < 0 ? -pb : pb, apc = pc < 0 ? -pc : pc;
        const int pred = (apa <= apb && apa <= apc) ? a[k] : (apb <= apc ? b : c[k]);
        a[k] = (cur[idx] + pred) & 255;
        c[k] = b;
        cur[idx] = (uint8_t)a[k];
    };
    for (; i + BPP <= rb; i += BPP)
        for (int k = 0; k < BPP; ++k) one(i + k, k);
    for (int k = 0; i + k < rb; ++k) one(i + k, k);
}
#define R3D_BPP_SWITCH(fn, bpp, generic, ...)      \
    switch (bpp) {                                 \
        case 1: fn<1>(__VA_ARGS__); break;         \
        case 2: fn<2>(__VA_ARGS__); break;         \
        case 3: fn<3>(__VA_ARGS__); break;         \
        case 4: fn<4>(__VA_ARGS__); break;         \
        case 6: fn<6>(__VA_ARGS__); break;         \
        case 8: fn<8>(__VA_ARGS__); break;         \
        default: generic;                          \
    }

inline void unfilter_sub(uint8_t* cur, size_t rb, size_t bpp) {
    switch (bpp) {
        case 1: unfilter_sub_n<1>(cur, rb); break;
        case 2: unfilter_sub_n<2>(cur, rb); break;
        case 3: unfilter_sub_n<3>(cur, rb); break;
        case 4: unfilter_sub_n<4>(cur, rb); break;
        case 6: unfilter_sub_n<6>(cur, rb); break;
        case 8: unfilter_sub_n<8>(cur, rb); break;
        default: for (size_t i = bpp; i < rb; ++i) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]);
    }
}

// inflate + un-filter in place: raw = H rows of (1 + rowbytes); afterwards row y's samples start at raw[y*(1+rb)+1]
bool inflate_unfilter(const PngHeader& h, const std::vector<uint8_t>& idat, std::vector<uint8_t>& raw, std::string& err) {
    const size_t bits = (size_t)h.channels() * h.depth;
    const size_t rb = ((size_t)h.W * bits + 7) / 8;
    const size_t bpp = bits >= 8 ? bits / 8 : 1;
    raw.resize((size_t)h.H * (rb + 1));
    if (r3d::inflate_zlib(idat.data(), idat.size(), raw.data(), raw.size()) != 0) {
        // anything the in-tree inflate does not take (it is strict about sizes and trailers) gets zlib's verdict
        uLongf dlen = (uLongf)raw.size();
        const int zr = uncompress(raw.data(), &dlen, idat.data(), (uLong)idat.size());
        if (zr != Z_OK || dlen != raw.size()) { err = "zlib inflate failed or image data has the wrong size"; return false; }
    }
    for (uint32_t y = 0; y < h.H; ++y) {
        uint8_t* row = &raw[(size_t)y * (rb + 1)];
        const int ft = row[0];
        uint8_t* cur = row + 1;
        const uint8_t* up = y ? cur - (rb + 1) : nullptr;
        switch (ft) {
            case 0: break;
            case 1: unfilter_sub(cur, rb, bpp); break;
            case 2: if (up) for (size_t i = 0; i < rb; ++i) cur[i] = (uint8_t)(cur[i] + up[i]); break;
            case 3:
                if (up) {
                    R3D_BPP_SWITCH(unfilter_avg_n, bpp, {
                        for (size_t i = 0; i < bpp && i < rb; ++i) cur[i] = (uint8_t)(cur[i] + (up[i] >> 1));
                        for (size_t i = bpp; i < rb; ++i) cur[i] = (uint8_t)(cur[i] + ((cur[i - bpp] + up[i]) >> 1));
                    }, cur, up, rb)
                } else {
                    for (size_t i = bpp; i < rb; ++i) cur[i] = (uint8_t)(cur[i] + (cur[i - bpp] >> 1));
                }
                break;
            case 4:
                if (up) {
                    R3D_BPP_SWITCH(unfilter_paeth_n, bpp, {
                        for (size_t i = 0; i < bpp && i < rb; ++i) cur[i] = (uint8_t)(cur[i] + up[i]);      // paeth(0, b, 0) = b
                        for (size_t i = bpp; i < rb; ++i) cur[i] = (uint8_t)(cur[i] + paeth(cur[i - bpp], up[i], up[i - bpp]));
                    }, cur, up, rb)
                } else {
                    for (size_t i = bpp; i < rb; ++i) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]);        // paeth(a, 0, 0) = a
                }
                break;
            default: err = "bad filter type"; return false;
        }
    }
    return true;
}

enum { MODE_GRAY8 = 0, MODE_CHANNEL = 1, MODE_RAW = 2 };

inline unsigned sample_sub8(const uint8_t* row, uint32_t x, int depth) {   // 1/2/4-bit sample x of a row
    const unsigned per = 8u / (unsigned)depth, idx = x / per, sh = (per - 1u - x % per) * (unsigned)depth;
    return (row[idx] >> sh) & ((1u << depth) - 1u);
}
inline unsigned gray8_from_rgb8(unsigned r, unsigned g, unsigned b) { return (9797u * r + 19234u * g + 3737u * b) >> 15; }
inline unsigned gray16_from_rgb16(unsigned r, unsigned g, unsigned b) { return (9797u * r + 19234u * g + 3737u * b + 16384u) >> 15; }

// one decoded frame -> the caller's W x H plane of elem_bytes-wide samples
bool convert(const PngHeader& h, const std::vector<uint8_t>& raw, const std::vector<uint8_t>& plte, int mode, int channel, void* out, int elem_bytes,
             std::string& err) {
    const int nc = h.channels();
    const size_t rb = ((size_t)h.W * nc * h.depth + 7) / 8;
    const int od = h.out_depth();
    if (mode == MODE_GRAY8 && elem_bytes != 1) { err = "GRAY8 output must be 1 byte per sample"; return false; }
    if (mode != MODE_GRAY8 && elem_bytes != od / 8) { err = "output sample width does not match the file's bit depth"; return false; }
    if (mode == MODE_CHANNEL) {
        const int cvc = h.cv_channels();
        if (cvc == 1) { err = "single-channel image has no channel axis (the reference's gt[:,:,1] raises IndexError)"; return false; }
        if (channel < 0 || channel >= cvc) { err = "channel index out of range"; return false; }
    }
    uint8_t* o8 = (uint8_t*)out;
    uint16_t* o16 = (uint16_t*)out;
    // fast paths for what depth / disparity maps are: single-channel 8 / 16 bit
    if (h.color == 0 && h.depth == 8 && mode != MODE_CHANNEL) {
        for (uint32_t y = 0; y < h.H; ++y) memcpy(o8 + (size_t)y * h.W, &raw[(size_t)y * (rb + 1) + 1], h.W);
        return true;
    }
    if (h.color == 0 && h.depth == 16 && mode == MODE_GRAY8) {
        for (uint32_t y = 0; y < h.H; ++y) {
            const uint8_t* row = &raw[(size_t)y * (rb + 1) + 1];
            uint8_t* o = o8 + (size_t)y * h.W;
            for (uint32_t x = 0; x < h.W; ++x) o[x] = row[2 * x];          // big-endian sample >> 8
        }
        return true;
    }
    if (h.color == 0 && h.depth == 16 && mode == MODE_RAW) {
        for (uint32_t y = 0; y < h.H; ++y) {
            const uint8_t* row = &raw[(size_t)y * (rb + 1) + 1];
            uint16_t* o = o16 + (size_t)y * h.W;
            for (uint32_t x = 0; x < h.W; ++x) o[x] = (uint16_t)(((unsigned)row[2 * x] << 8) | row[2 * x + 1]);
        }
        return true;
    }
    for (uint32_t y = 0; y < h.H; ++y) {
        const uint8_t* row = &raw[(size_t)y * (rb + 1) + 1];
        const size_t ob = (size_t)y * h.W;
        for (uint32_t x = 0; x < h.W; ++x) {
            // fetch the pixel as (r, g, b, a, grey) at the file's depth (8 or 16), palette / sub-byte grey expanded
            unsigned r = 0, g = 0, b = 0, a = od == 16 ? 65535u : 255u, grey = 0;
            bool is_grey = false;
            if (h.color == 0) {
                is_grey = true;
                if (h.depth == 16) grey = ((unsigned)row[2 * x] << 8) | row[2 * x + 1];
                else if (h.depth == 8) grey = row[x];
                else grey = sample_sub8(row, x, h.depth) * (255u / ((1u << h.depth) - 1u));
            } else if (h.color == 4) {
                is_grey = true;
                if (h.depth == 16) { grey = ((unsigned)row[4 * x] << 8) | row[4 * x + 1]; a = ((unsigned)row[4 * x + 2] << 8) | row[4 * x + 3]; }
                else { grey = row[2 * x]; a = row[2 * x + 1]; }
            } else if (h.color == 3) {
                const unsigned idx = h.depth == 8 ? row[x] : sample_sub8(row, x, h.depth);
                if (3 * idx + 2 < plte.size()) { r = plte[3 * idx]; g = plte[3 * idx + 1]; b = plte[3 * idx + 2]; }
            } else {
                const int step = (h.color == 6 ? 4 : 3) * (h.depth / 8);
                const uint8_t* p = row + (size_t)x * step;
                if (h.depth == 16) {
                    r = ((unsigned)p[0] << 8) | p[1]; g = ((unsigned)p[2] << 8) | p[3]; b = ((unsigned)p[4] << 8) | p[5];
                    if (h.color == 6) a = ((unsigned)p[6] << 8) | p[7];
                } else {
                    r = p[0]; g = p[1]; b = p[2];
                    if (h.color == 6) a = p[3];
                }
            }
            unsigned v;
            if (mode == MODE_GRAY8) {
                if (is_grey) v = od == 16 ? grey >> 8 : grey;
                else v = od == 16 ? gray16_from_rgb16(r, g, b) >> 8 : gray8_from_rgb8(r, g, b);
                o8[ob + x] = (uint8_t)v;
                continue;
            }
            if (mode == MODE_RAW && h.cv_channels() == 1) v = grey;
            else {
                const int c = mode == MODE_RAW ? 0 : channel;
                if (is_grey) v = c == 3 ? a : grey;                  // grey+alpha expands to B = G = R = grey, A
                else v = c == 0 ? b : (c == 1 ? g : (c == 2 ? r : a));
            }
            if (od == 16) o16[ob + x] = (uint16_t)v; else o8[ob + x] = (uint8_t)v;
        }
    }
    return true;
}

bool decode_one(const char* path, int mode, int channel, void* out, int elem_bytes, int W, int H, std::string& err) {
    // per-thread buffers, kept between files (a fresh 1 MB vector is an mmap + page faults + zero fill every time)
    static thread_local std::vector<uint8_t> file, idat, plte, raw;
    idat.clear();
    plte.clear();
    PngHeader h;
    bool trns = false;
    if (!read_file(path, file, err)) return false;
    if (!parse(file, h, idat, plte, trns, err)) return false;
    if ((int)h.W != W || (int)h.H != H) { char b[128]; snprintf(b, sizeof b, "image is %ux%u, expected %dx%d", h.W, h.H, W, H); err = b; return false; }
    if (trns && mode != MODE_GRAY8) { err = "tRNS transparency with IMREAD_UNCHANGED is not supported"; return false; }
    if (!inflate_unfilter(h, idat, raw, err)) return false;
    return convert(h, raw, plte, mode, channel, out, elem_bytes, err);
}

}  // namespace

// Header of a PNG: width, height, channels as IMREAD_UNCHANGED would shape it (1 = no channel axis, 3, 4), bits per
// sample after expansion (8 or 16).
extern "C" int r3d_png_info(const char* path, int* W, int* H, int* channels, int* bit_depth) {
    if (!path) return set_error(nullptr, R3D_ERR_ARG, "null path");
    std::vector<uint8_t> file, idat, plte;
    std::string err;
    PngHeader h;
    bool trns = false;
    if (!read_file(path, file, err)) return set_error(nullptr, R3D_ERR_IO, "%s", err.c_str());
    if (!parse(file, h, idat, plte, trns, err)) return set_error(nullptr, R3D_ERR_UNSUPPORTED, "%s: %s", path, err.c_str());
    if (W) *W = (int)h.W;
    if (H) *H = (int)h.H;
    if (channels) *channels = h.cv_channels();
    if (bit_depth) *bit_depth = h.out_depth();
    return R3D_OK;
}

extern "C" int r3d_png_decode_batch(const char* const* paths, int n, int mode, int channel, void* out, size_t frame_stride_bytes, int elem_bytes,
                                    int W, int H, int n_threads, int* status) {
    if (n < 0 || (n && (!paths || !out))) return set_error(nullptr, R3D_ERR_ARG, "r3d_png_decode_batch: null argument");
    if (mode < MODE_GRAY8 || mode > MODE_RAW) return set_error(nullptr, R3D_ERR_ARG, "bad decode mode %d", mode);
    if (W <= 0 || H <= 0 || (elem_bytes != 1 && elem_bytes != 2)) return set_error(nullptr, R3D_ERR_ARG, "bad frame geometry");
    if (frame_stride_bytes == 0) frame_stride_bytes = (size_t)W * H * elem_bytes;
    if (frame_stride_bytes < (size_t)W * H * elem_bytes) return set_error(nullptr, R3D_ERR_ARG, "frame stride smaller than a frame");
    if (n == 0) return R3D_OK;
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    int nt = n_threads > 0 ? n_threads : (int)hw;
    if (nt > n) nt = n;
    std::atomic<int> next(0), failed(0);
    std::vector<std::string> errs((size_t)nt);
    std::vector<int> first_bad((size_t)nt, -1);
    auto work = [&](int tid) {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n) break;
            std::string err;
            const bool ok = paths[i] && decode_one(paths[i], mode, channel, (char*)out + (size_t)i * frame_stride_bytes, elem_bytes, W, H, err);
            if (status) status[i] = ok ? R3D_OK : R3D_ERR_IO;
            if (!ok) {
                failed.fetch_add(1);
                if (first_bad[tid] < 0) { first_bad[tid] = i; errs[tid] = (paths[i] ? std::string(paths[i]) : std::string("(null)")) + ": " + err; }
            }
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; ++t) pool.emplace_back(work, t);
        for (auto& th : pool) th.join();
    }
    if (failed.load()) {
        int best = -1;
        for (int t = 0; t < nt; ++t) if (first_bad[t] >= 0 && (best < 0 || first_bad[t] < first_bad[best])) best = t;
        return set_error(nullptr, R3D_ERR_IO, "%d of %d PNG files failed; first: %s", failed.load(), n, best >= 0 ? errs[best].c_str() : "?");
    }
    return R3D_OK;
}
