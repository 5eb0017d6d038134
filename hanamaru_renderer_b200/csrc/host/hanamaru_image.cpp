// hanamaru_image.cpp -- the host's image I/O: what the reference gets from the `image` crate (image 0.19.0,
// src/texture.rs:16-20 `image::open`, src/renderer.rs:97 / src/main.rs:1217 `save`).
//   png_encode_rgb8   result.png / NNN.png (8-bit RGB, zlib deflate, filter 0 / Paeth chosen per row)
//   png_decode        8 / 16-bit, grey / grey+alpha / RGB / RGBA / palette, non-interlaced
//   jpeg_decode       baseline sequential DCT (SOF0), Huffman, 8-bit, 1 or 3 components, any integral sampling factors
//                     with "fancy" triangle upsampling for h2v1 / h2v2 -- every JPEG under the reference's textures/
//                     is SOF0 4:2:0
// The crates are not in the tree (Cargo.lock pins png 0.11 / jpeg-decoder 0.1.15): these are restatements of the
// published formats.  PNG is lossless, so any correct decoder gives the reference's texels.  JPEG is not: the inverse
// DCT, the chroma upsampling filter and the colour conversion are implementation choices.  This decoder uses the
// integer IDCT (12-bit constants, the public-domain stb_image formulation that jpeg-decoder's idct.rs follows), libjpeg's
// fancy upsampling and the floating-point BT.601 conversion; tests bound its distance to libjpeg-turbo (PIL) per texel.
// Decoded RGBA8 texels are INPUTS of the C ABI (hnm_image), shared by the CUDA core and the oracle.
#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "hanamaru_host.h"

namespace hanamaru {

// ------------------------------------------------------------------------------------------------ PNG
static void put_be32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}
static void put_chunk(std::vector<uint8_t>& out, const char* type, const uint8_t* data, size_t n) {
    put_be32(out, (uint32_t)n);
    size_t at = out.size();
    out.insert(out.end(), type, type + 4);
    if (n) out.insert(out.end(), data, data + n);
    put_be32(out, (uint32_t)crc32(0L, out.data() + at, (uInt)(n + 4)));
}
static inline int paeth(int a, int b, int c) {
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

bool png_encode_rgb8(const uint8_t* rgb, uint32_t width, uint32_t height, std::vector<uint8_t>& out, std::string* err) {
    out.clear();
    if (!rgb || width == 0 || height == 0) { if (err) *err = "png_encode: empty image"; return false; }
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    out.insert(out.end(), sig, sig + 8);
    uint8_t ihdr[13];
    ihdr[0] = (uint8_t)(width >> 24); ihdr[1] = (uint8_t)(width >> 16); ihdr[2] = (uint8_t)(width >> 8); ihdr[3] = (uint8_t)width;
    ihdr[4] = (uint8_t)(height >> 24); ihdr[5] = (uint8_t)(height >> 16); ihdr[6] = (uint8_t)(height >> 8); ihdr[7] = (uint8_t)height;
    ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;  // 8-bit RGB, deflate, adaptive filtering, no interlace
    put_chunk(out, "IHDR", ihdr, 13);
    const size_t stride = (size_t)width * 3;
    std::vector<uint8_t> raw((stride + 1) * height);
    std::vector<uint8_t> cand(stride);
    for (uint32_t y = 0; y < height; y++) {
        const uint8_t* row = rgb + stride * y;
        const uint8_t* up = y ? rgb + stride * (y - 1) : nullptr;
        // two candidates per row (None, Paeth), the smaller sum of absolute values wins (the usual heuristic)
        uint64_t s_none = 0, s_paeth = 0;
        for (size_t i = 0; i < stride; i++) {
            int a = i >= 3 ? row[i - 3] : 0, b = up ? up[i] : 0, c = (up && i >= 3) ? up[i - 3] : 0;
            uint8_t f = (uint8_t)(row[i] - paeth(a, b, c));
            cand[i] = f;
            s_paeth += (uint64_t)std::abs((int)(int8_t)f);
            s_none += (uint64_t)std::abs((int)(int8_t)row[i]);
        }
        uint8_t* dst = raw.data() + (stride + 1) * y;
        if (s_paeth < s_none) { dst[0] = 4; memcpy(dst + 1, cand.data(), stride); }
        else { dst[0] = 0; memcpy(dst + 1, row, stride); }
    }
    uLongf bound = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(bound);
    if (compress2(z.data(), &bound, raw.data(), (uLong)raw.size(), 6) != Z_OK) { if (err) *err = "png_encode: deflate failed"; return false; }
    put_chunk(out, "IDAT", z.data(), bound);
    put_chunk(out, "IEND", nullptr, 0);
    return true;
}

static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

bool png_decode(const uint8_t* data, size_t n, Image& out, std::string* err) {
    auto fail = [&](const char* m) { if (err) *err = std::string("png_decode: ") + m; return false; };
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    if (n < 8 || memcmp(data, sig, 8) != 0) return fail("not a PNG");
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    size_t at = 8;
    bool seen_ihdr = false, seen_iend = false;
    while (at + 12 <= n && !seen_iend) {
        uint32_t len = be32(data + at);
        const uint8_t* type = data + at + 4;
        const uint8_t* body = data + at + 8;
        if (at + 12 + (size_t)len > n) return fail("truncated chunk");
        if (be32(body + len) != (uint32_t)crc32(0L, type, (uInt)(len + 4))) return fail("chunk CRC mismatch");
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) return fail("bad IHDR");
            w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
            if (body[10] != 0 || body[11] != 0) return fail("unknown compression / filter method");
            seen_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) plte.assign(body, body + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(body, body + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(type, "IEND", 4)) seen_iend = true;
        at += 12 + (size_t)len;
    }
    if (!seen_ihdr || w == 0 || h == 0) return fail("no IHDR");
    if (interlace != 0) return fail("Adam7 interlacing is not supported");
    int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!channels) return fail("bad colour type");
    if (!(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4)))) return fail("bad bit depth");
    const size_t bpp_bits = (size_t)channels * depth;
    const size_t stride = ((size_t)w * bpp_bits + 7) / 8;
    const size_t fb = std::max<size_t>(1, bpp_bits / 8);  // filter unit in bytes
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawn = (uLongf)raw.size();
    int zr = uncompress(raw.data(), &rawn, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || rawn != raw.size()) return fail("inflate failed");
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    out.width = w; out.height = h;
    out.rgba.assign((size_t)w * h * 4, 255);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t* src = raw.data() + (stride + 1) * y;
        const int ft = src[0];
        for (size_t i = 0; i < stride; i++) {
            int a = i >= fb ? cur[i - fb] : 0, b = prev[i], c = i >= fb ? prev[i - fb] : 0;
            int x = src[1 + i];
            switch (ft) {
                case 0: break;
                case 1: x += a; break;
                case 2: x += b; break;
                case 3: x += (a + b) >> 1; break;
                case 4: x += paeth(a, b, c); break;
                default: return fail("bad filter type");
            }
            cur[i] = (uint8_t)x;
        }
        uint8_t* dst = out.rgba.data() + (size_t)w * 4 * y;
        for (uint32_t x = 0; x < w; x++) {
            auto sample = [&](int ch) -> uint32_t {  // 8-bit value of channel ch of pixel x (16-bit: the high byte, as image 0.19 scales)
                if (depth == 16) return cur[((size_t)x * channels + ch) * 2];
                if (depth == 8) return cur[(size_t)x * channels + ch];
                const size_t bit = (size_t)x * depth;
                const uint32_t v = (cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
                return ctype == 3 ? v : v * 255u / ((1u << depth) - 1u);
            };
            uint8_t r, g, b_, a = 255;
            if (ctype == 3) {
                uint32_t idx = sample(0);
                if ((size_t)idx * 3 + 2 >= plte.size()) return fail("palette index out of range");
                r = plte[idx * 3]; g = plte[idx * 3 + 1]; b_ = plte[idx * 3 + 2];
                if (idx < trns.size()) a = trns[idx];
            } else if (ctype == 0 || ctype == 4) {
                r = g = b_ = (uint8_t)sample(0);
                if (ctype == 4) a = (uint8_t)sample(1);
            } else {
                r = (uint8_t)sample(0); g = (uint8_t)sample(1); b_ = (uint8_t)sample(2);
                if (ctype == 6) a = (uint8_t)sample(3);
            }
            dst[4 * x] = r; dst[4 * x + 1] = g; dst[4 * x + 2] = b_; dst[4 * x + 3] = a;
        }
        std::swap(prev, cur);
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ JPEG (baseline)
namespace {
struct Huff {
    // canonical code lookup: for code length L (1..16), codes [mincode[L], maxcode[L]] map to vals[valptr[L] + code - mincode[L]]
    int mincode[17], maxcode[18], valptr[17];
    uint8_t vals[256];
    bool present = false;
};
struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint32_t acc = 0;
    int nbits = 0;
    bool hit_marker = false;
    void fill() {
        while (nbits <= 24) {
            int b = 0;
            if (!hit_marker && p < end) {
                b = *p++;
                if (b == 0xFF) {
                    int b2 = p < end ? *p : 0;
                    if (b2 == 0) p++;            // stuffed zero
                    else { hit_marker = true; p--; b = 0; }  // a marker: feed zeros from here on
                }
            }
            acc |= (uint32_t)b << (24 - nbits);
            nbits += 8;
        }
    }
    int bit() {
        if (nbits < 1) fill();
        int b = (int)(acc >> 31);
        acc <<= 1; nbits--;
        return b;
    }
    int bits(int n) {
        if (n == 0) return 0;
        if (nbits < n) fill();
        int v = (int)(acc >> (32 - n));
        acc <<= n; nbits -= n;
        return v;
    }
    void reset() { acc = 0; nbits = 0; hit_marker = false; }
};
int huff_decode(BitReader& br, const Huff& h) {
    int code = 0;
    for (int len = 1; len <= 16; len++) {
        code = (code << 1) | br.bit();
        if (h.maxcode[len] >= 0 && code <= h.maxcode[len] && code >= h.mincode[len]) return h.vals[h.valptr[len] + code - h.mincode[len]];
    }
    return -1;
}
inline int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }
const uint8_t ZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                            41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                            30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
inline uint8_t clamp8(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }
// integer inverse DCT with 12-bit constants (the public-domain stb_image formulation)
#define IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)                                      \
    int t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3;                          \
    p2 = s2; p3 = s6;                                                                \
    p1 = (p2 + p3) * 2217;              /* 0.5411961   */                            \
    t2 = p1 + p3 * -7567;               /* -1.847759065 */                           \
    t3 = p1 + p2 * 3135;                /* 0.765366865 */                            \
    p2 = s0; p3 = s4;                                                                \
    t0 = (p2 + p3) * 4096; t1 = (p2 - p3) * 4096;                                    \
    x0 = t0 + t3; x3 = t0 - t3; x1 = t1 + t2; x2 = t1 - t2;                          \
    t0 = s7; t1 = s5; t2 = s3; t3 = s1;                                              \
    p3 = t0 + t2; p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2;                          \
    p5 = (p3 + p4) * 4816;              /* 1.175875602 */                            \
    t0 = t0 * 1223;                     /* 0.298631336 */                            \
    t1 = t1 * 8410;                     /* 2.053119869 */                            \
    t2 = t2 * 12586;                    /* 3.072711026 */                            \
    t3 = t3 * 6149;                     /* 1.501321110 */                            \
    p1 = p5 + p1 * -3685;               /* -0.899976223 */                           \
    p2 = p5 + p2 * -10497;              /* -2.562915447 */                           \
    p3 = p3 * -8034;                    /* -1.961570560 */                           \
    p4 = p4 * -1597;                    /* -0.390180644 */                           \
    t3 += p1 + p4; t2 += p2 + p3; t1 += p2 + p4; t0 += p1 + p3;
void idct_block(const int* d, uint8_t* out, int out_stride) {
    int val[64];
    for (int i = 0; i < 8; i++) {
        const int* s = d + i;
        int* v = val + i;
        if (s[8] == 0 && s[16] == 0 && s[24] == 0 && s[32] == 0 && s[40] == 0 && s[48] == 0 && s[56] == 0) {
            int dc = s[0] * 4;
            v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dc;
        } else {
            IDCT_1D(s[0], s[8], s[16], s[24], s[32], s[40], s[48], s[56])
            x0 += 512; x1 += 512; x2 += 512; x3 += 512;
            v[0] = (x0 + t3) >> 10; v[56] = (x0 - t3) >> 10;
            v[8] = (x1 + t2) >> 10; v[48] = (x1 - t2) >> 10;
            v[16] = (x2 + t1) >> 10; v[40] = (x2 - t1) >> 10;
            v[24] = (x3 + t0) >> 10; v[32] = (x3 - t0) >> 10;
        }
    }
    for (int i = 0; i < 8; i++) {
        const int* v = val + 8 * i;
        uint8_t* o = out + (size_t)out_stride * i;
        IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
        x0 += 65536 + (128 << 17); x1 += 65536 + (128 << 17); x2 += 65536 + (128 << 17); x3 += 65536 + (128 << 17);
        o[0] = clamp8((x0 + t3) >> 17); o[7] = clamp8((x0 - t3) >> 17);
        o[1] = clamp8((x1 + t2) >> 17); o[6] = clamp8((x1 - t2) >> 17);
        o[2] = clamp8((x2 + t1) >> 17); o[5] = clamp8((x2 - t1) >> 17);
        o[3] = clamp8((x3 + t0) >> 17); o[4] = clamp8((x3 - t0) >> 17);
    }
}
struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int bw = 0, bh = 0;  // blocks per line / column (padded to whole MCUs)
    std::vector<uint8_t> plane;
    int pred = 0;
};
}  // namespace

bool jpeg_decode(const uint8_t* data, size_t n, Image& out, std::string* err) {
    auto fail = [&](const char* m) { if (err) *err = std::string("jpeg_decode: ") + m; return false; };
    if (n < 4 || data[0] != 0xFF || data[1] != 0xD8) return fail("not a JPEG");
    uint16_t qt[4][64];
    bool qt_ok[4] = {false, false, false, false};
    Huff hdc[4], hac[4];
    std::vector<Component> comps;
    int width = 0, height = 0, hmax = 1, vmax = 1, restart_interval = 0;
    size_t at = 2;
    bool done = false, have_frame = false;
    while (!done) {
        while (at < n && data[at] != 0xFF) at++;
        while (at < n && data[at] == 0xFF) at++;
        if (at >= n) break;
        const int marker = data[at++];
        if (marker == 0xD9) break;
        if (marker == 0x01 || (marker >= 0xD0 && marker <= 0xD7)) continue;
        if (at + 2 > n) return fail("truncated segment");
        const size_t len = ((size_t)data[at] << 8) | data[at + 1];
        if (len < 2 || at + len > n) return fail("bad segment length");
        const uint8_t* seg = data + at + 2;
        const size_t sl = len - 2;
        switch (marker) {
            case 0xDB: {  // DQT
                size_t i = 0;
                while (i < sl) {
                    int pq = seg[i] >> 4, tq = seg[i] & 15;
                    i++;
                    if (tq > 3) return fail("bad quantisation table id");
                    for (int k = 0; k < 64; k++) {
                        if (i + (pq ? 2 : 1) > sl) return fail("truncated DQT");
                        qt[tq][ZIGZAG[k]] = pq ? (uint16_t)((seg[i] << 8) | seg[i + 1]) : seg[i];
                        i += pq ? 2 : 1;
                    }
                    qt_ok[tq] = true;
                }
                break;
            }
            case 0xC4: {  // DHT
                size_t i = 0;
                while (i + 17 <= sl) {
                    int tc = seg[i] >> 4, th = seg[i] & 15;
                    if (th > 3 || tc > 1) return fail("bad Huffman table id");
                    Huff& h = tc ? hac[th] : hdc[th];
                    int counts[17];
                    int total = 0;
                    for (int l = 1; l <= 16; l++) { counts[l] = seg[i + l]; total += counts[l]; }
                    i += 17;
                    if (total > 256 || i + (size_t)total > sl) return fail("bad DHT");
                    memcpy(h.vals, seg + i, (size_t)total);
                    i += (size_t)total;
                    int code = 0, k = 0;
                    for (int l = 1; l <= 16; l++) {
                        h.valptr[l] = k; h.mincode[l] = code;
                        code += counts[l]; k += counts[l];
                        h.maxcode[l] = counts[l] ? code - 1 : -1;
                        code <<= 1;
                    }
                    h.present = true;
                }
                break;
            }
            case 0xC0: case 0xC1: {  // SOF0 / SOF1: sequential Huffman
                if (sl < 6 || seg[0] != 8) return fail("only 8-bit precision is supported");
                height = (seg[1] << 8) | seg[2]; width = (seg[3] << 8) | seg[4];
                int nc = seg[5];
                if ((nc != 1 && nc != 3) || sl < 6 + 3 * (size_t)nc || width == 0 || height == 0) return fail("unsupported frame header");
                comps.resize((size_t)nc);
                for (int c = 0; c < nc; c++) {
                    comps[c].id = seg[6 + 3 * c]; comps[c].h = seg[7 + 3 * c] >> 4; comps[c].v = seg[7 + 3 * c] & 15; comps[c].tq = seg[8 + 3 * c];
                    if (comps[c].h < 1 || comps[c].h > 4 || comps[c].v < 1 || comps[c].v > 4 || comps[c].tq > 3) return fail("bad component");
                    hmax = std::max(hmax, comps[c].h); vmax = std::max(vmax, comps[c].v);
                }
                have_frame = true;
                break;
            }
            case 0xC2: return fail("progressive JPEG is not supported");
            case 0xDD: if (sl >= 2) restart_interval = (seg[0] << 8) | seg[1]; break;
            case 0xDA: {  // SOS + entropy-coded data
                if (!have_frame) return fail("scan before frame");
                int ns = seg[0];
                if (ns != (int)comps.size()) return fail("only single-scan (interleaved) baseline files are supported");
                for (int k = 0; k < ns; k++) {
                    int cid = seg[1 + 2 * k];
                    bool found = false;
                    for (auto& c : comps) if (c.id == cid) { c.td = seg[2 + 2 * k] >> 4; c.ta = seg[2 + 2 * k] & 15; found = true; }
                    if (!found) return fail("scan names an unknown component");
                }
                const int mcux = (width + 8 * hmax - 1) / (8 * hmax), mcuy = (height + 8 * vmax - 1) / (8 * vmax);
                for (auto& c : comps) {
                    c.bw = mcux * c.h; c.bh = mcuy * c.v;
                    c.plane.assign((size_t)c.bw * 8 * c.bh * 8, 0);
                    c.pred = 0;
                    if (!qt_ok[c.tq] || !hdc[c.td].present || !hac[c.ta].present) return fail("missing table");
                }
                BitReader br{data + at + len, data + n};
                int restarts_left = restart_interval;
                int coef[64];
                for (int my = 0; my < mcuy; my++) {
                    for (int mx = 0; mx < mcux; mx++) {
                        if (restart_interval && restarts_left == 0) {
                            // byte-align, expect RSTn
                            br.reset();
                            const uint8_t* q = br.p;
                            while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) q++;
                            if (q + 1 >= br.end) return fail("missing restart marker");
                            br.p = q + 2;
                            for (auto& c : comps) c.pred = 0;
                            restarts_left = restart_interval;
                        }
                        for (auto& c : comps) {
                            for (int by = 0; by < c.v; by++) {
                                for (int bx = 0; bx < c.h; bx++) {
                                    memset(coef, 0, sizeof(coef));
                                    int t = huff_decode(br, hdc[c.td]);
                                    if (t < 0 || t > 11) return fail("bad DC code");
                                    int diff = t ? extend(br.bits(t), t) : 0;
                                    c.pred += diff;
                                    coef[0] = c.pred * qt[c.tq][0];
                                    for (int k = 1; k < 64;) {
                                        int rs = huff_decode(br, hac[c.ta]);
                                        if (rs < 0) return fail("bad AC code");
                                        int r = rs >> 4, s = rs & 15;
                                        if (s == 0) {
                                            if (r == 15) { k += 16; continue; }
                                            break;  // EOB
                                        }
                                        k += r;
                                        if (k > 63) return fail("AC run past the block");
                                        coef[ZIGZAG[k]] = extend(br.bits(s), s) * qt[c.tq][ZIGZAG[k]];
                                        k++;
                                    }
                                    const size_t px = ((size_t)mx * c.h + bx) * 8, py = ((size_t)my * c.v + by) * 8;
                                    idct_block(coef, c.plane.data() + py * ((size_t)c.bw * 8) + px, c.bw * 8);
                                }
                            }
                        }
                        if (restart_interval) restarts_left--;
                    }
                }
                done = true;
                break;
            }
            default: break;  // APPn, COM, ...
        }
        at += len;
    }
    if (!done) return fail("no scan found");
    out.width = (uint32_t)width; out.height = (uint32_t)height;
    out.rgba.assign((size_t)width * height * 4, 255);
    // chroma planes at full resolution
    std::vector<std::vector<uint8_t>> full(comps.size());
    for (size_t ci = 0; ci < comps.size(); ci++) {
        const Component& c = comps[ci];
        const int sw = c.bw * 8;                                  // stored width of the plane
        const int cw = (width * c.h + hmax - 1) / hmax, ch = (height * c.v + vmax - 1) / vmax;  // meaningful samples
        const int fx = hmax / c.h, fy = vmax / c.v;
        std::vector<uint8_t>& f = full[ci];
        f.resize((size_t)width * height);
        if (fx == 1 && fy == 1) {
            for (int y = 0; y < height; y++) memcpy(f.data() + (size_t)y * width, c.plane.data() + (size_t)y * sw, (size_t)width);
        } else if (fx == 2 && (fy == 2 || fy == 1) && hmax % c.h == 0 && vmax % c.v == 0) {
            // libjpeg's "fancy" upsampling: triangle filter, 3/4 near + 1/4 far per axis
            std::vector<int> sum((size_t)cw);
            for (int y = 0; y < height; y++) {
                const int sy = fy == 2 ? y >> 1 : y;
                int far_y = fy == 2 ? ((y & 1) ? sy + 1 : sy - 1) : sy;
                far_y = far_y < 0 ? 0 : (far_y >= ch ? ch - 1 : far_y);
                const uint8_t* near_row = c.plane.data() + (size_t)std::min(sy, ch - 1) * sw;
                const uint8_t* far_row = c.plane.data() + (size_t)far_y * sw;
                uint8_t* o = f.data() + (size_t)y * width;
                if (fy == 2) {
                    for (int x = 0; x < cw; x++) sum[x] = 3 * near_row[x] + far_row[x];
                    for (int x = 0; x < cw; x++) {
                        const int cur = sum[x], prev = x ? sum[x - 1] : cur, next = x + 1 < cw ? sum[x + 1] : cur;
                        const int a = x ? (3 * cur + prev + 8) >> 4 : (4 * cur + 8) >> 4;
                        const int b = x + 1 < cw ? (3 * cur + next + 7) >> 4 : (4 * cur + 7) >> 4;
                        if (2 * x < width) o[2 * x] = (uint8_t)a;
                        if (2 * x + 1 < width) o[2 * x + 1] = (uint8_t)b;
                    }
                } else {
                    for (int x = 0; x < cw; x++) {
                        const int cur = near_row[x], prev = x ? near_row[x - 1] : cur, next = x + 1 < cw ? near_row[x + 1] : cur;
                        const int a = x ? (3 * cur + prev + 1) >> 2 : cur;
                        const int b = x + 1 < cw ? (3 * cur + next + 2) >> 2 : cur;
                        if (2 * x < width) o[2 * x] = (uint8_t)a;
                        if (2 * x + 1 < width) o[2 * x + 1] = (uint8_t)b;
                    }
                }
            }
        } else {
            if (hmax % c.h != 0 || vmax % c.v != 0) return fail("fractional sampling ratios are not supported");
            for (int y = 0; y < height; y++)
                for (int x = 0; x < width; x++) f[(size_t)y * width + x] = c.plane[(size_t)std::min(y / fy, ch - 1) * sw + std::min(x / fx, cw - 1)];
        }
    }
    for (size_t i = 0; i < (size_t)width * height; i++) {
        uint8_t* px = out.rgba.data() + 4 * i;
        if (comps.size() == 1) { px[0] = px[1] = px[2] = full[0][i]; continue; }
        // ITU-R BT.601, full range, evaluated in f32 and rounded half up
        const float y = (float)full[0][i], cb = (float)full[1][i] - 128.0f, cr = (float)full[2][i] - 128.0f;
        const float r = y + 1.40200f * cr, g = y - 0.34414f * cb - 0.71414f * cr, b = y + 1.77200f * cb;
        px[0] = clamp8((int)(r + 0.5f)); px[1] = clamp8((int)(g + 0.5f)); px[2] = clamp8((int)(b + 0.5f));
    }
    return true;
}

// `image::open` (src/texture.rs:18): the format follows the file's magic bytes
bool image_decode(const uint8_t* data, size_t n, Image& out, std::string* err) {
    if (n >= 8 && data[0] == 0x89 && data[1] == 'P') return png_decode(data, n, out, err);
    if (n >= 2 && data[0] == 0xFF && data[1] == 0xD8) return jpeg_decode(data, n, out, err);
    if (err) *err = "image_decode: neither PNG nor JPEG";
    return false;
}

}  // namespace hanamaru
