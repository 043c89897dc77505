// fpm_jpeg.h -- host half of the JPEG ingest (the reference reads its sources with
// cv::imread(path, IMREAD_GRAYSCALE), src/MatchToolDialog.cpp:314, :341, and two of its Test Images are JPEG files).
//
// What cv::imread does for a JPEG in grayscale mode (OpenCV grfmt_jpeg.cpp over libjpeg-turbo): out_color_space =
// JCS_GRAYSCALE, i.e. only the first (luma) component is reconstructed, with the default accurate integer IDCT (JDCT_ISLOW,
// jidctint.c) -- no chroma, no upsampling, no colour conversion.  That makes the decode exactly reproducible:
//   host  : marker parsing + Huffman decoding of the entropy-coded segment (inherently sequential) -> the quantised luma
//           coefficients of every 8x8 block (int16, natural order) + the luma quantisation table          [this file]
//   device: dequantisation + ISLOW IDCT + range limit, one thread per block -> the u8 frame in HBM         [fpm_kernels.cuh]
// Supported: baseline / extended-sequential Huffman JPEG (SOF0 / SOF1), 8-bit, one scan, grayscale or YCbCr with any chroma
// subsampling (luma at full resolution), restart intervals.  Rejected with a message: progressive, arithmetic, lossless,
// 12-bit, CMYK / RGB-coded files, multi-scan files.
#pragma once
#include <stdint.h>
#include <string.h>
#include <new>
#include <string>
#include <vector>

namespace fpm_jpeg {

struct Luma {
    int width = 0, height = 0;         // image size
    int bw = 0, bh = 0;                // luma blocks per row / column (whole MCUs)
    uint16_t quant[64];                // luma quantisation table, natural (row-major) order
    std::vector<int16_t> coef;         // [bh * bw][64] quantised coefficients, natural order
};

static const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
    bool defined = false;
    std::vector<uint16_t> look;        // 16-bit prefix -> (length << 8) | symbol, 0 = invalid code
    // AC tables: 16-bit prefix -> code AND value bits in one step when both fit the window:
    // (value + 2048) << 12 | run << 8 | total bits (code + size); 0 = take the two-step path
    std::vector<uint32_t> fast;
    void build_fast()
    {
        fast.assign(65536, 0);
        for (uint32_t w = 0; w < 65536; w++) {
            const uint32_t e = look[w];
            if (!e) continue;
            const int len = e >> 8, r = (e >> 4) & 15, sz = e & 15;
            if (sz == 0 || sz > 11 || len + sz > 16) continue;      // |value| <= 2047: value + 2048 stays a positive 12-bit field
            int v = (int)((w >> (16 - len - sz)) & ((1u << sz) - 1));
            if (v < (1 << (sz - 1))) v = v - (1 << sz) + 1;
            fast[w] = ((uint32_t)(v + 2048) << 12) | ((uint32_t)r << 8) | (uint32_t)(len + sz);
        }
    }
    void build(const uint8_t counts[16], const uint8_t* symbols)
    {
        look.assign(65536, 0);
        uint32_t code = 0;
        int k = 0;
        for (int len = 1; len <= 16; len++) {
            for (int i = 0; i < counts[len - 1]; i++, k++, code++) {
                if (code >= (1u << len)) return;                         // over-subscribed table: leave the rest invalid
                const uint32_t first = code << (16 - len), n = 1u << (16 - len);
                for (uint32_t j = 0; j < n; j++) look[first + j] = (uint16_t)((len << 8) | symbols[k]);
            }
            code <<= 1;
        }
        defined = true;
    }
};

// bit reader over the entropy-coded segment: removes the 0xFF00 stuffing, stops at a marker (then feeds zero bits)
struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t acc = 0;                  // bits left-aligned
    int nbits = 0;
    int marker = 0;                    // marker that ended the data (0 = none yet)
    BitReader(const uint8_t* b, const uint8_t* e) : p(b), end(e) {}
    void fill()
    {
        // four bytes at once while none of them is 0xFF (no stuffing, no marker): the common case
        while (nbits <= 32 && !marker && p + 4 <= end) {
            uint32_t v;
            memcpy(&v, p, 4);
            if (((v & 0x7F7F7F7Fu) + 0x01010101u) & v & 0x80808080u) break;
            v = __builtin_bswap32(v);
            acc |= (uint64_t)v << (32 - nbits);
            nbits += 32;
            p += 4;
        }
        while (nbits <= 56) {
            uint32_t byte = 0;
            if (!marker && p < end) {
                if (*p == 0xFF) {
                    if (p + 1 < end && p[1] == 0x00) { byte = 0xFF; p += 2; }
                    else if (p + 1 < end && p[1] == 0xFF) { p++; continue; }          // fill byte
                    else { marker = p + 1 < end ? p[1] : 0xD9; }
                } else {
                    byte = *p++;
                }
            }
            acc |= (uint64_t)byte << (56 - nbits);
            nbits += 8;
        }
    }
    uint32_t peek16() { if (nbits < 16) fill(); return (uint32_t)(acc >> 48); }
    void skip(int n) { acc <<= n; nbits -= n; }
    int receive(int n)
    {
        if (n == 0) return 0;
        if (nbits < n) fill();
        const int v = (int)(acc >> (64 - n));
        skip(n);
        return v;
    }
    // after a restart interval: drop the partial byte, consume the RSTn marker
    bool restart()
    {
        acc = 0; nbits = 0;
        if (!marker) {                                                   // the marker has not been seen by fill() yet
            while (p + 1 < end && !(p[0] == 0xFF && p[1] != 0x00 && p[1] != 0xFF)) p++;
            if (p + 1 >= end) return false;
            marker = p[1];
        }
        if (marker < 0xD0 || marker > 0xD7) return false;
        p += 2;
        marker = 0;
        return true;
    }
};

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

struct Component { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0; };

inline uint32_t be16(const uint8_t* p) { return ((uint32_t)p[0] << 8) | p[1]; }

// everything the headers say, up to the first byte of the entropy-coded segment
struct Frame {
    int width = 0, height = 0;
    std::vector<Component> comp;       // frame order; comp[0] = luma
    int hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
    int restart_interval = 0;
    uint16_t qt[4][64];                // natural order
    bool qt_def[4] = {false, false, false, false};
    uint8_t dc_counts[4][16], ac_counts[4][16], dc_syms[4][256], ac_syms[4][256];   // the DHT specifications as they came
    HuffTable dc[4], ac[4];            // lookup tables of the sequential host decoder
    size_t scan_begin = 0;             // offset of the entropy-coded data
};

// returns "" on success, else what is wrong / unsupported
inline std::string parse(const uint8_t* d, size_t n, Frame* f)
{
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return "not a JPEG file (no SOI)";
    uint16_t (&qt)[4][64] = f->qt;
    bool (&qt_def)[4] = f->qt_def;
    HuffTable (&dc)[4] = f->dc;
    HuffTable (&ac)[4] = f->ac;
    std::vector<Component>& comp = f->comp;
    int restart_interval = 0, width = 0, height = 0;
    bool saw_jfif = false, saw_adobe = false, sof = false;
    int adobe_transform = 0;
    size_t i = 2;
    while (true) {
        while (i < n && d[i] != 0xFF) i++;                               // tolerate garbage between segments, like libjpeg
        while (i < n && d[i] == 0xFF) i++;
        if (i >= n) return "truncated JPEG (no SOS)";
        const int m = d[i++];
        if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (m == 0xD9) return "JPEG without image data (EOI before SOS)";
        if (i + 2 > n) return "truncated JPEG segment";
        const size_t len = be16(d + i);
        if (len < 2 || i + len > n) return "truncated JPEG segment";
        const uint8_t* s = d + i + 2;
        const size_t sl = len - 2;
        if (m == 0xDB) {                                                 // DQT
            size_t k = 0;
            while (k < sl) {
                const int pq = s[k] >> 4, tq = s[k] & 15;
                k++;
                if (tq > 3 || pq > 1 || k + (pq ? 128 : 64) > sl) return "bad DQT segment";
                for (int z = 0; z < 64; z++) { qt[tq][kZigzag[z]] = pq ? (uint16_t)be16(s + k + 2 * z) : s[k + z]; }
                k += pq ? 128 : 64;
                qt_def[tq] = true;
            }
        } else if (m == 0xC4) {                                          // DHT
            size_t k = 0;
            while (k < sl) {
                const int tc = s[k] >> 4, th = s[k] & 15;
                k++;
                if (tc > 1 || th > 3 || k + 16 > sl) return "bad DHT segment";
                int total = 0;
                for (int z = 0; z < 16; z++) total += s[k + z];
                if (total > 256 || k + 16 + total > sl) return "bad DHT segment";
                (tc ? ac[th] : dc[th]).defined = true;                     // the 64 K-entry lookup tables are built by the host decoder only
                memcpy(tc ? f->ac_counts[th] : f->dc_counts[th], s + k, 16);
                memcpy(tc ? f->ac_syms[th] : f->dc_syms[th], s + k + 16, total);
                k += 16 + total;
            }
        } else if (m == 0xC0 || m == 0xC1) {                             // SOF0 / SOF1
            if (sof) return "JPEG with two frame headers";
            if (sl < 6) return "bad SOF segment";
            if (s[0] != 8) return "unsupported JPEG: " + std::to_string((int)s[0]) + "-bit samples";
            height = (int)be16(s + 1); width = (int)be16(s + 3);
            const int nf = s[5];
            if (height == 0 || width == 0) return "unsupported JPEG: zero size / DNL";
            if (nf != 1 && nf != 3) return "unsupported JPEG: " + std::to_string(nf) + " components";
            if (sl < (size_t)(6 + 3 * nf)) return "bad SOF segment";
            comp.resize(nf);
            for (int c = 0; c < nf; c++) {
                comp[c].id = s[6 + 3 * c]; comp[c].h = s[7 + 3 * c] >> 4; comp[c].v = s[7 + 3 * c] & 15; comp[c].tq = s[8 + 3 * c];
                if (comp[c].h < 1 || comp[c].h > 4 || comp[c].v < 1 || comp[c].v > 4 || comp[c].tq > 3) return "bad SOF segment";
            }
            sof = true;
        } else if (m == 0xC2 || m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
            return std::string("unsupported JPEG: ") + (m == 0xC2 ? "progressive" : m >= 0xC9 ? "arithmetic coding" : "lossless / hierarchical");
        } else if (m == 0xDD) {                                          // DRI
            if (sl < 2) return "bad DRI segment";
            restart_interval = (int)be16(s);
        } else if (m == 0xE0) {
            if (sl >= 5 && memcmp(s, "JFIF\0", 5) == 0) saw_jfif = true;
        } else if (m == 0xEE) {
            if (sl >= 12 && memcmp(s, "Adobe", 5) == 0) { saw_adobe = true; adobe_transform = s[11]; }
        } else if (m == 0xDA) {                                          // SOS
            if (!sof) return "JPEG scan before the frame header";
            if (sl < 1) return "bad SOS segment";
            const int ns = s[0];
            if (ns != (int)comp.size()) return "unsupported JPEG: more than one scan";
            if (sl < (size_t)(1 + 2 * ns + 3)) return "bad SOS segment";
            for (int c = 0; c < ns; c++) {
                if (s[1 + 2 * c] != comp[c].id) return "unsupported JPEG: scan components out of frame order";
                comp[c].td = s[2 + 2 * c] >> 4; comp[c].ta = s[2 + 2 * c] & 15;
                if (comp[c].td > 3 || comp[c].ta > 3 || !dc[comp[c].td].defined || !ac[comp[c].ta].defined) return "JPEG scan uses an undefined Huffman table";
            }
            if (s[1 + 2 * ns] != 0 || s[2 + 2 * ns] != 63 || s[3 + 2 * ns] != 0) return "unsupported JPEG: spectral selection / successive approximation";
            i += len;
            break;
        }
        i += len;
    }
    const int nf = (int)comp.size();
    if (nf == 3) {
        // libjpeg's colour-space guess (jdapimin.c default_decompress_parms): the grayscale output of an RGB-coded file is a
        // weighted sum of three planes, not the first component -- not supported here
        bool ycc = true;
        if (saw_jfif) ycc = true;
        else if (saw_adobe) ycc = adobe_transform != 0;
        else if (comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B') ycc = false;
        if (!ycc) return "unsupported JPEG: RGB-coded components";
    }
    int hmax = 1, vmax = 1;
    for (auto& c : comp) { hmax = c.h > hmax ? c.h : hmax; vmax = c.v > vmax ? c.v : vmax; }
    if (nf == 1) { comp[0].h = comp[0].v = 1; hmax = vmax = 1; }          // a non-interleaved scan: one block per MCU
    if (comp[0].h != hmax || comp[0].v != vmax) return "unsupported JPEG: subsampled luma";
    if (!qt_def[comp[0].tq]) return "JPEG frame uses an undefined quantisation table";
    const int mcux = (width + 8 * hmax - 1) / (8 * hmax), mcuy = (height + 8 * vmax - 1) / (8 * vmax);
    f->width = width; f->height = height; f->hmax = hmax; f->vmax = vmax; f->mcux = mcux; f->mcuy = mcuy;
    f->restart_interval = restart_interval;
    f->scan_begin = i;
    return "";
}

// the sequential decoder: Huffman decoding of the whole scan on one host thread
inline std::string decode_scan_host(const Frame& fr, const uint8_t* d, size_t n, Luma* out)
{
    const std::vector<Component>& comp = fr.comp;
    HuffTable dc[4], ac[4];
    for (const Component& c : comp) {
        if (dc[c.td].look.empty()) dc[c.td].build(fr.dc_counts[c.td], fr.dc_syms[c.td]);
        if (ac[c.ta].look.empty()) { ac[c.ta].build(fr.ac_counts[c.ta], fr.ac_syms[c.ta]); ac[c.ta].build_fast(); }
    }
    const int nf = (int)comp.size(), mcux = fr.mcux, mcuy = fr.mcuy, restart_interval = fr.restart_interval;
    const size_t i = fr.scan_begin;
    out->width = fr.width; out->height = fr.height;
    out->bw = mcux * comp[0].h; out->bh = mcuy * comp[0].v;
    memcpy(out->quant, fr.qt[comp[0].tq], sizeof(out->quant));
    try {
        out->coef.assign((size_t)out->bw * out->bh * 64, 0);
    } catch (const std::bad_alloc&) {
        return "JPEG too large for host memory";
    }

    BitReader br(d + i, d + n);
    int pred[3] = {0, 0, 0};
    int16_t scratch[64];
    int todo = restart_interval;
    for (int my = 0; my < mcuy; my++) {
        for (int mx = 0; mx < mcux; mx++) {
            if (restart_interval && todo == 0) {
                if (!br.restart()) return "corrupt JPEG: restart marker missing";
                pred[0] = pred[1] = pred[2] = 0;
                todo = restart_interval;
            }
            todo--;
            for (int c = 0; c < nf; c++) {
                const HuffTable& hd = dc[comp[c].td];
                const HuffTable& ha = ac[comp[c].ta];
                for (int by = 0; by < comp[c].v; by++) {
                    for (int bx = 0; bx < comp[c].h; bx++) {
                        int16_t* blk = c == 0 ? &out->coef[((size_t)(my * comp[0].v + by) * out->bw + (mx * comp[0].h + bx)) * 64] : scratch;
                        uint32_t e = hd.look[br.peek16()];
                        if (!e) return "corrupt JPEG: bad Huffman code";
                        br.skip(e >> 8);
                        int sz = e & 255;
                        if (sz > 15) return "corrupt JPEG: bad DC size";
                        if (sz) pred[c] += extend(br.receive(sz), sz);
                        if (c == 0) blk[0] = (int16_t)pred[c];
                        for (int k = 1; k < 64;) {
                            const uint32_t w16 = br.peek16();
                            const uint32_t f = ha.fast[w16];
                            if (f) {                                     // code + value bits inside the 16-bit window
                                br.skip(f & 255);
                                k += (f >> 8) & 15;
                                if (k > 63) return "corrupt JPEG: coefficient index out of range";
                                if (c == 0) blk[kZigzag[k]] = (int16_t)((int)(f >> 12) - 2048);
                                k++;
                                continue;
                            }
                            e = ha.look[w16];
                            if (!e) return "corrupt JPEG: bad Huffman code";
                            br.skip(e >> 8);
                            const int r = (e >> 4) & 15;
                            sz = e & 15;
                            if (sz == 0) {
                                if (r != 15) break;                      // EOB
                                k += 16;
                                continue;
                            }
                            k += r;
                            const int v = extend(br.receive(sz), sz);
                            if (k > 63) return "corrupt JPEG: coefficient index out of range";
                            if (c == 0) blk[kZigzag[k]] = (int16_t)v;
                            k++;
                        }
                    }
                }
            }
        }
    }
    return "";
}

inline std::string decode_luma(const uint8_t* d, size_t n, Luma* out)
{
    Frame fr;
    const std::string why = parse(d, n, &fr);
    if (!why.empty()) return why;
    return decode_scan_host(fr, d, n, out);
}

}  // namespace fpm_jpeg
