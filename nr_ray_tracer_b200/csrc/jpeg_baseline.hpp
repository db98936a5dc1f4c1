// jpeg_baseline.hpp — minimal baseline (SOF0, 8-bit, Huffman) JPEG decoder for scene textures.
//
// The reference decodes textures with image 0.25.8 / zune-jpeg (textures/image.rs:24) into 8-bit RGB and then
// into_rgb32f() (u8/255).  No JPEG library headers exist in the build image, so this is a from-scratch decoder of
// the subset the shipped textures use (scenes/textures/{earth,moon}.jpg: SOF0, 3 components, any sampling
// factors up to 2x2, optional restart intervals, Adobe APP14 transform flag).  It follows the JPEG standard's
// decoding procedure with the "slow-but-accurate" integer inverse DCT of Loeffler, Ligtenberg & Moschytz
// (13-bit constants, two passes) and 16-bit fixed-point YCbCr->RGB, which is what libjpeg-class decoders
// produce bit for bit (checked against PIL in tests/test_native_loader.py).  Parity with zune-jpeg itself is
// unpinned (decoders may differ by <= 1 LSB per texel).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace nrrt_jpeg {

struct Image {
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> rgb;  // row-major, 3 bytes per pixel
};

struct Huff {
    uint8_t bits[17] = {0};
    uint8_t vals[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    bool present = false;
    void build() {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; ++l) {
            valptr[l] = k;
            mincode[l] = code;
            code += bits[l];
            k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        present = true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int dc_pred = 0;
    int blocks_w = 0, blocks_h = 0;  // padded to whole MCUs
    std::vector<uint8_t> plane;      // blocks_w*8 x blocks_h*8 samples
};

class Decoder {
  public:
    bool decode(const uint8_t* data, size_t size, Image& out, std::string& err) {
        p_ = data;
        end_ = data + size;
        if (size < 4 || p_[0] != 0xFF || p_[1] != 0xD8) return fail(err, "not a JPEG (no SOI)");
        p_ += 2;
        bool have_sof = false;
        for (;;) {
            if (p_ + 4 > end_) return fail(err, "truncated JPEG");
            if (*p_ != 0xFF) return fail(err, "marker expected");
            while (p_ < end_ && *p_ == 0xFF) ++p_;
            if (p_ >= end_) return fail(err, "truncated JPEG");
            uint8_t m = *p_++;
            if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
            if (m == 0xD9) return fail(err, "EOI before image data");
            if (p_ + 2 > end_) return fail(err, "truncated JPEG");
            size_t len = ((size_t)p_[0] << 8) | p_[1];
            if (len < 2 || p_ + len > end_) return fail(err, "bad segment length");
            const uint8_t* seg = p_ + 2;
            size_t n = len - 2;
            switch (m) {
                case 0xDB:
                    if (!read_dqt(seg, n)) return fail(err, "bad DQT");
                    break;
                case 0xC4:
                    if (!read_dht(seg, n)) return fail(err, "bad DHT");
                    break;
                case 0xC0:
                case 0xC1:
                    if (!read_sof(seg, n, err)) return false;
                    have_sof = true;
                    break;
                case 0xC2:
                    return fail(err, "progressive JPEG is not supported (baseline only)");
                case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
                    return fail(err, "unsupported JPEG coding process");
                case 0xDD:
                    if (n < 2) return fail(err, "bad DRI");
                    restart_interval_ = (seg[0] << 8) | seg[1];
                    break;
                case 0xEE:  // Adobe APP14: transform flag in byte 11
                    if (n >= 12 && std::memcmp(seg, "Adobe", 5) == 0) {
                        adobe_ = true;
                        adobe_transform_ = seg[11];
                    }
                    break;
                case 0xDA: {
                    if (!have_sof) return fail(err, "SOS before SOF");
                    if (!read_sos(seg, n)) return fail(err, "bad SOS");
                    p_ += len;
                    if (!decode_scan(err)) return false;
                    return finish(out, err);
                }
                default: break;  // APPn, COM: skipped
            }
            p_ += len;
        }
    }

  private:
    const uint8_t *p_ = nullptr, *end_ = nullptr;
    uint16_t qt_[4][64] = {{0}};
    bool qt_present_[4] = {false, false, false, false};
    Huff dc_[4], ac_[4];
    std::vector<Component> comps_;
    int width_ = 0, height_ = 0, hmax_ = 1, vmax_ = 1, restart_interval_ = 0;
    bool adobe_ = false;
    int adobe_transform_ = 0;
    // bit reader
    uint32_t bitbuf_ = 0;
    int bitcnt_ = 0;
    bool hit_marker_ = false;

    static bool fail(std::string& err, const char* m) {
        err = m;
        return false;
    }
    static const uint8_t* zigzag() {
        static const uint8_t z[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                      41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                      30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
        return z;
    }
    bool read_dqt(const uint8_t* s, size_t n) {
        while (n > 0) {
            int pq = s[0] >> 4, tq = s[0] & 15;
            if (tq > 3) return false;
            size_t need = 1 + (pq ? 128 : 64);
            if (n < need) return false;
            for (int i = 0; i < 64; ++i)
                qt_[tq][zigzag()[i]] = pq ? (uint16_t)((s[1 + 2 * i] << 8) | s[2 + 2 * i]) : s[1 + i];
            qt_present_[tq] = true;
            s += need;
            n -= need;
        }
        return true;
    }
    bool read_dht(const uint8_t* s, size_t n) {
        while (n > 0) {
            if (n < 17) return false;
            int tc = s[0] >> 4, th = s[0] & 15;
            if (tc > 1 || th > 3) return false;
            Huff& h = tc ? ac_[th] : dc_[th];
            int total = 0;
            for (int l = 1; l <= 16; ++l) total += (h.bits[l] = s[l]);
            if (total > 256 || n < (size_t)(17 + total)) return false;
            std::memcpy(h.vals, s + 17, total);
            h.build();
            s += 17 + total;
            n -= 17 + total;
        }
        return true;
    }
    bool read_sof(const uint8_t* s, size_t n, std::string& err) {
        if (n < 6) return fail(err, "bad SOF");
        if (s[0] != 8) return fail(err, "only 8-bit JPEG is supported");
        height_ = (s[1] << 8) | s[2];
        width_ = (s[3] << 8) | s[4];
        int nc = s[5];
        if (width_ <= 0 || height_ <= 0 || (nc != 1 && nc != 3) || n < (size_t)(6 + 3 * nc)) return fail(err, "bad SOF");
        comps_.resize(nc);
        for (int i = 0; i < nc; ++i) {
            comps_[i].id = s[6 + 3 * i];
            comps_[i].h = s[7 + 3 * i] >> 4;
            comps_[i].v = s[7 + 3 * i] & 15;
            comps_[i].tq = s[8 + 3 * i];
            if (comps_[i].h < 1 || comps_[i].h > 2 || comps_[i].v < 1 || comps_[i].v > 2 || comps_[i].tq > 3)
                return fail(err, "unsupported sampling factors");
            hmax_ = comps_[i].h > hmax_ ? comps_[i].h : hmax_;
            vmax_ = comps_[i].v > vmax_ ? comps_[i].v : vmax_;
        }
        return true;
    }
    bool read_sos(const uint8_t* s, size_t n) {
        if (n < 1) return false;
        int ns = s[0];
        if (ns != (int)comps_.size() || n < (size_t)(1 + 2 * ns + 3)) return false;  // interleaved single scan only
        for (int i = 0; i < ns; ++i) {
            int id = s[1 + 2 * i];
            bool found = false;
            for (auto& c : comps_)
                if (c.id == id) {
                    c.td = s[2 + 2 * i] >> 4;
                    c.ta = s[2 + 2 * i] & 15;
                    if (c.td > 3 || c.ta > 3) return false;
                    found = true;
                }
            if (!found) return false;
        }
        return true;
    }
    // ---- entropy decoding
    inline void fill() {
        while (bitcnt_ <= 24) {
            uint32_t b = 0;
            if (!hit_marker_ && p_ < end_) {
                b = *p_;
                if (b == 0xFF) {
                    if (p_ + 1 < end_ && p_[1] == 0x00) {
                        p_ += 2;
                    } else {
                        hit_marker_ = true;  // leave the marker in place; feed zeros
                        b = 0;
                    }
                } else {
                    ++p_;
                }
            }
            bitbuf_ |= b << (24 - bitcnt_);
            bitcnt_ += 8;
        }
    }
    inline int get_bits(int n) {
        if (n == 0) return 0;
        if (bitcnt_ < n) fill();
        int v = (int)(bitbuf_ >> (32 - n));
        bitbuf_ <<= n;
        bitcnt_ -= n;
        return v;
    }
    inline int decode_huff(const Huff& h) {
        if (bitcnt_ < 16) fill();
        int code = 0;
        for (int l = 1; l <= 16; ++l) {
            code = (code << 1) | (int)(bitbuf_ >> 31);
            bitbuf_ <<= 1;
            --bitcnt_;
            if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l])
                return h.vals[h.valptr[l] + code - h.mincode[l]];
        }
        return -1;
    }
    static inline int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }

    bool decode_block(Component& c, int16_t* blk) {
        std::memset(blk, 0, 64 * sizeof(int16_t));
        const Huff& hd = dc_[c.td];
        const Huff& ha = ac_[c.ta];
        int t = decode_huff(hd);
        if (t < 0 || t > 11) return false;
        int diff = t ? extend(get_bits(t), t) : 0;
        c.dc_pred += diff;
        blk[0] = (int16_t)c.dc_pred;
        for (int k = 1; k < 64;) {
            int rs = decode_huff(ha);
            if (rs < 0) return false;
            int r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (r == 15) {
                    k += 16;
                    continue;
                }
                break;  // EOB
            }
            k += r;
            if (k > 63) return false;
            blk[zigzag()[k]] = (int16_t)extend(get_bits(s), s);
            ++k;
        }
        return true;
    }

    // "islow" integer IDCT (Loeffler-Ligtenberg-Moschytz), CONST_BITS = 13, PASS1_BITS = 2
    static inline int descale(int64_t x, int n) { return (int)((x + ((int64_t)1 << (n - 1))) >> n); }
    static void idct(const int16_t* in, const uint16_t* q, uint8_t* out, int stride) {
        const int C0_298 = 2446, C0_390 = 3196, C0_541 = 4433, C0_765 = 6270, C0_899 = 7373, C1_175 = 9633,
                  C1_501 = 12299, C1_847 = 15137, C1_961 = 16069, C2_053 = 16819, C2_562 = 20995, C3_072 = 25172;
        int ws[64];
        for (int col = 0; col < 8; ++col) {
            int d[8];
            for (int r = 0; r < 8; ++r) d[r] = in[r * 8 + col] * (int)q[r * 8 + col];
            if (!(d[1] | d[2] | d[3] | d[4] | d[5] | d[6] | d[7])) {
                int dc = d[0] * 4;  // multiplications, not shifts: the operands may be negative
                for (int r = 0; r < 8; ++r) ws[r * 8 + col] = dc;
                continue;
            }
            int64_t z2 = d[2], z3 = d[6];
            int64_t z1 = (z2 + z3) * C0_541;
            int64_t tmp2 = z1 + z3 * (-C1_847), tmp3 = z1 + z2 * C0_765;
            z2 = d[0], z3 = d[4];
            int64_t tmp0 = (z2 + z3) * 8192, tmp1 = (z2 - z3) * 8192;
            int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            tmp0 = d[7], tmp1 = d[5], tmp2 = d[3], tmp3 = d[1];
            z1 = tmp0 + tmp3, z2 = tmp1 + tmp2, z3 = tmp0 + tmp2;
            int64_t z4 = tmp1 + tmp3, z5 = (z3 + z4) * C1_175;
            tmp0 *= C0_298, tmp1 *= C2_053, tmp2 *= C3_072, tmp3 *= C1_501;
            z1 *= -C0_899, z2 *= -C2_562, z3 *= -C1_961, z4 *= -C0_390;
            z3 += z5, z4 += z5;
            tmp0 += z1 + z3, tmp1 += z2 + z4, tmp2 += z2 + z3, tmp3 += z1 + z4;
            ws[0 * 8 + col] = descale(tmp10 + tmp3, 11);
            ws[7 * 8 + col] = descale(tmp10 - tmp3, 11);
            ws[1 * 8 + col] = descale(tmp11 + tmp2, 11);
            ws[6 * 8 + col] = descale(tmp11 - tmp2, 11);
            ws[2 * 8 + col] = descale(tmp12 + tmp1, 11);
            ws[5 * 8 + col] = descale(tmp12 - tmp1, 11);
            ws[3 * 8 + col] = descale(tmp13 + tmp0, 11);
            ws[4 * 8 + col] = descale(tmp13 - tmp0, 11);
        }
        for (int row = 0; row < 8; ++row) {
            const int* w = ws + row * 8;
            uint8_t* o = out + row * stride;
            int64_t z2 = w[2], z3 = w[6];
            int64_t z1 = (z2 + z3) * C0_541;
            int64_t tmp2 = z1 + z3 * (-C1_847), tmp3 = z1 + z2 * C0_765;
            int64_t tmp0 = ((int64_t)w[0] + w[4]) * 8192, tmp1 = ((int64_t)w[0] - w[4]) * 8192;
            int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            tmp0 = w[7], tmp1 = w[5], tmp2 = w[3], tmp3 = w[1];
            z1 = tmp0 + tmp3, z2 = tmp1 + tmp2, z3 = tmp0 + tmp2;
            int64_t z4 = tmp1 + tmp3, z5 = (z3 + z4) * C1_175;
            tmp0 *= C0_298, tmp1 *= C2_053, tmp2 *= C3_072, tmp3 *= C1_501;
            z1 *= -C0_899, z2 *= -C2_562, z3 *= -C1_961, z4 *= -C0_390;
            z3 += z5, z4 += z5;
            tmp0 += z1 + z3, tmp1 += z2 + z4, tmp2 += z2 + z3, tmp3 += z1 + z4;
            auto put = [](int64_t v) { int x = descale(v, 18) + 128; return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); };
            o[0] = put(tmp10 + tmp3), o[7] = put(tmp10 - tmp3), o[1] = put(tmp11 + tmp2), o[6] = put(tmp11 - tmp2);
            o[2] = put(tmp12 + tmp1), o[5] = put(tmp12 - tmp1), o[3] = put(tmp13 + tmp0), o[4] = put(tmp13 - tmp0);
        }
    }

    bool decode_scan(std::string& err) {
        for (auto& c : comps_) {
            if (!qt_present_[c.tq] || !dc_[c.td].present || !ac_[c.ta].present) return fail(err, "missing JPEG table");
        }
        const int mcu_w = 8 * hmax_, mcu_h = 8 * vmax_;
        const int mcus_x = (width_ + mcu_w - 1) / mcu_w, mcus_y = (height_ + mcu_h - 1) / mcu_h;
        for (auto& c : comps_) {
            c.blocks_w = mcus_x * c.h;
            c.blocks_h = mcus_y * c.v;
            c.plane.assign((size_t)c.blocks_w * 8 * c.blocks_h * 8, 0);
            c.dc_pred = 0;
        }
        bitbuf_ = 0, bitcnt_ = 0, hit_marker_ = false;
        int16_t blk[64];
        int until_restart = restart_interval_;
        for (int my = 0; my < mcus_y; ++my)
            for (int mx = 0; mx < mcus_x; ++mx) {
                if (restart_interval_ && until_restart == 0) {
                    // byte-align, expect RSTn
                    bitbuf_ = 0, bitcnt_ = 0, hit_marker_ = false;
                    while (p_ + 1 < end_ && !(p_[0] == 0xFF && p_[1] >= 0xD0 && p_[1] <= 0xD7)) ++p_;
                    if (p_ + 1 >= end_) return fail(err, "missing restart marker");
                    p_ += 2;
                    for (auto& c : comps_) c.dc_pred = 0;
                    until_restart = restart_interval_;
                }
                for (auto& c : comps_)
                    for (int by = 0; by < c.v; ++by)
                        for (int bx = 0; bx < c.h; ++bx) {
                            if (!decode_block(c, blk)) return fail(err, "corrupt JPEG entropy data");
                            int stride = c.blocks_w * 8;
                            uint8_t* dst = c.plane.data() + (size_t)((my * c.v + by) * 8) * stride + (mx * c.h + bx) * 8;
                            idct(blk, qt_[c.tq], dst, stride);
                        }
                if (restart_interval_) --until_restart;
            }
        return true;
    }

    bool finish(Image& out, std::string& err) {
        out.width = (uint32_t)width_;
        out.height = (uint32_t)height_;
        out.rgb.resize((size_t)width_ * height_ * 3);
        if (comps_.size() == 1) {
            const Component& c = comps_[0];
            for (int y = 0; y < height_; ++y)
                for (int x = 0; x < width_; ++x) {
                    uint8_t v = c.plane[(size_t)y * c.blocks_w * 8 + x];
                    uint8_t* o = &out.rgb[((size_t)y * width_ + x) * 3];
                    o[0] = o[1] = o[2] = v;
                }
            return true;
        }
        // upsample chroma by replication when subsampled ("fancy" upsampling is not reproduced; 4:4:4 needs none)
        const bool ycc = adobe_ ? (adobe_transform_ != 0) : true;
        // libjpeg's fixed-point tables (SCALEBITS = 16)
        const int ONE_HALF = 1 << 15;
        auto FIX = [](double v) { return (int)(v * 65536.0 + 0.5); };
        for (int y = 0; y < height_; ++y)
            for (int x = 0; x < width_; ++x) {
                int s[3];
                for (int k = 0; k < 3; ++k) {
                    const Component& c = comps_[k];
                    int sx = x * c.h / hmax_, sy = y * c.v / vmax_;
                    s[k] = c.plane[(size_t)sy * c.blocks_w * 8 + sx];
                }
                uint8_t* o = &out.rgb[((size_t)y * width_ + x) * 3];
                if (!ycc) {
                    o[0] = (uint8_t)s[0], o[1] = (uint8_t)s[1], o[2] = (uint8_t)s[2];
                    continue;
                }
                int yy = s[0], cb = s[1] - 128, cr = s[2] - 128;
                int r = yy + ((FIX(1.40200) * cr + ONE_HALF) >> 16);
                int g = yy + ((-FIX(0.34414) * cb + ONE_HALF - FIX(0.71414) * cr) >> 16);
                int b = yy + ((FIX(1.77200) * cb + ONE_HALF) >> 16);
                o[0] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
                o[1] = (uint8_t)(g < 0 ? 0 : (g > 255 ? 255 : g));
                o[2] = (uint8_t)(b < 0 ? 0 : (b > 255 ? 255 : b));
            }
        (void)err;
        return true;
    }
};

inline bool decode(const uint8_t* data, size_t size, Image& out, std::string& err) {
    Decoder d;
    return d.decode(data, size, out, err);
}

}  // namespace nrrt_jpeg
