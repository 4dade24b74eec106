// jpeg_decode.cpp -- JPEG ingest for the drop-in driver (baseline, extended-sequential and
// progressive Huffman JPEG, 8-bit precision, restart intervals), written from ITU-T T.81.
//
// Why an own decoder: the decoder is part of the reference's contract.  Its drivers call
// stbi_load(name, &W, &H, &nCh, 1) (Deff2D.cuh:342, cuh:377) and threshold the returned bytes at
// 150 / 200 / 50 (cuh:1779, cuh:1456-1467); decoders that differ by one grey level move pixels
// across those thresholds (364 pixels of the bundled 00042.jpg in 2-phase
// mode).  For single-component (grayscale) files -- the only ones the reference accepts,
// cuh:1665-1668 -- a decoder's output is fixed by the entropy decoding (exact by the standard),
// the dequantisation, and the inverse DCT.  This file uses the same inverse DCT as the
// reference's decoder: the Loeffler-Ligtenberg-Moschytz "islow" integer transform with 12-bit
// constants, a column pass keeping 2 extra bits and a row pass that rounds, removes the level
// shift and clamps in one step, on 16-bit dequantised coefficients.  tests/test_host_logic.py
// checks it pixel for pixel against the reference's own decoder on the bundled images and on
// generated baseline / progressive / restart-interval files.
//
// Colour files are decoded far enough to report their component count (the drivers reject
// them); their gray output is the luma plane, not the reference's RGB-derived luma.
#include "deff2d_internal.h"

#include <cstring>

namespace deff2d {

namespace {

const uint8_t kZigzag[64 + 15] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                  41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                  30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
                                  // run-off protection for corrupt streams
                                  63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

struct Huff {
    // canonical Huffman code (T.81 Annex C / F.2.2.3): per length the smallest code, the index of
    // its value and the largest code
    int mincode[17] = {0}, maxcode[18] = {0}, valptr[17] = {0};
    uint8_t vals[256] = {0};
    bool present = false;

    bool build(const uint8_t *bits, const uint8_t *v, int n)
    {
        std::memcpy(vals, v, (size_t)n);
        int code = 0, k = 0;
        for (int len = 1; len <= 16; len++) {
            valptr[len] = k;
            mincode[len] = code;
            code += bits[len - 1];
            k += bits[len - 1];
            maxcode[len] = bits[len - 1] ? code - 1 : -1;
            if (code > (1 << len)) return false;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        present = true;
        return true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0;
    int td = 0, ta = 0;           // Huffman table selectors of the current scan
    int bw = 0, bh = 0;           // blocks per row / column, padded to whole MCUs
    int cw = 0, ch = 0;           // blocks per row / column covering the component's own samples
    int dc_pred = 0;
    std::vector<int16_t> coef;    // progressive: bw * bh * 64 coefficients
    std::vector<uint8_t> pix;     // bw*8 x bh*8 samples
};

struct Decoder {
    const uint8_t *d;
    size_t n, pos = 0;
    std::string *err;
    int W = 0, H = 0, ncomp = 0;
    bool progressive = false;
    uint16_t qt[4][64];
    bool qt_present[4] = {false, false, false, false};
    Huff hdc[4], hac[4];
    Component comp[4];
    int hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
    int restart_interval = 0;
    // bit reader
    uint32_t bitbuf = 0;
    int bitcnt = 0;
    int marker = 0;               // marker met inside entropy-coded data (0: none)
    bool nomore = false;
    // progressive scan state
    int ss = 0, se = 0, ah = 0, al = 0, eobrun = 0;

    bool fail(const char *m) { *err = std::string("JPEG: ") + m; return false; }

    int u8() { return pos < n ? d[pos++] : 0; }
    int u16() { const int a = u8(); return (a << 8) | u8(); }

    // ---- entropy-coded segment bit reader (T.81 F.2.2.5: 0xFF00 stuffing, markers end the data)
    void fill()
    {
        while (bitcnt <= 24) {
            int b = 0;
            if (!nomore) {
                if (pos >= n) nomore = true;
                else {
                    b = d[pos++];
                    if (b == 0xFF) {
                        int c = pos < n ? d[pos++] : 0xD9;
                        while (c == 0xFF) c = pos < n ? d[pos++] : 0xD9;     // fill bytes
                        if (c != 0) { marker = c; nomore = true; b = 0; }   // a marker ends the entropy-coded data
                    }
                }
            }
            bitbuf |= (uint32_t)b << (24 - bitcnt);
            bitcnt += 8;
        }
    }
    int getbits(int k)
    {
        if (k == 0) return 0;
        if (bitcnt < k) fill();
        const int v = (int)(bitbuf >> (32 - k));
        bitbuf <<= k;
        bitcnt -= k;
        return v;
    }
    int getbit() { return getbits(1); }
    int decode(const Huff &h)
    {
        if (bitcnt < 16) fill();
        int code = 0;
        for (int len = 1; len <= 16; len++) {
            code = (int)(bitbuf >> (32 - len));
            if (h.maxcode[len] >= 0 && code <= h.maxcode[len] && code >= h.mincode[len]) {
                bitbuf <<= len;
                bitcnt -= len;
                return h.vals[h.valptr[len] + code - h.mincode[len]];
            }
        }
        return -1;
    }
    // T.81 F.2.2.1 EXTEND
    int receive_extend(int s)
    {
        if (s == 0) return 0;
        const int v = getbits(s);
        return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
    }
    void reset_entropy()
    {
        bitbuf = 0; bitcnt = 0; marker = 0; nomore = false; eobrun = 0;
        for (int c = 0; c < ncomp; c++) comp[c].dc_pred = 0;
    }

    // ---- segments -------------------------------------------------------------------------
    bool read_dqt(int len)
    {
        while (len > 0) {
            const int pq = u8(), prec = pq >> 4, t = pq & 15;
            if (t > 3 || prec > 1) return fail("bad DQT");
            for (int k = 0; k < 64; k++) qt[t][kZigzag[k]] = (uint16_t)(prec ? u16() : u8());
            qt_present[t] = true;
            len -= 1 + (prec ? 128 : 64);
        }
        return len == 0 || fail("bad DQT length");
    }
    bool read_dht(int len)
    {
        while (len > 0) {
            const int tc = u8(), cls = tc >> 4, t = tc & 15;
            if (cls > 1 || t > 3) return fail("bad DHT");
            uint8_t bits[16], vals[256];
            int total = 0;
            for (int k = 0; k < 16; k++) { bits[k] = (uint8_t)u8(); total += bits[k]; }
            if (total > 256) return fail("bad DHT");
            for (int k = 0; k < total; k++) vals[k] = (uint8_t)u8();
            if (!(cls ? hac[t] : hdc[t]).build(bits, vals, total)) return fail("bad Huffman code lengths");
            len -= 17 + total;
        }
        return len == 0 || fail("bad DHT length");
    }
    bool read_sof(int len)
    {
        const int prec = u8();
        H = u16();
        W = u16();
        ncomp = u8();
        if (prec != 8) return fail("only 8-bit precision is supported");
        if (W < 1 || H < 1) return fail("empty image");
        if (ncomp != 1 && ncomp != 3 && ncomp != 4) return fail("bad component count");
        if (len != 6 + 3 * ncomp) return fail("bad SOF length");
        hmax = vmax = 1;
        for (int c = 0; c < ncomp; c++) {
            comp[c].id = u8();
            const int hv = u8();
            comp[c].h = hv >> 4; comp[c].v = hv & 15; comp[c].tq = u8();
            if (comp[c].h < 1 || comp[c].h > 4 || comp[c].v < 1 || comp[c].v > 4 || comp[c].tq > 3) return fail("bad SOF component");
            if (comp[c].h > hmax) hmax = comp[c].h;
            if (comp[c].v > vmax) vmax = comp[c].v;
        }
        mcux = (W + 8 * hmax - 1) / (8 * hmax);
        mcuy = (H + 8 * vmax - 1) / (8 * vmax);
        for (int c = 0; c < ncomp; c++) {
            Component &k = comp[c];
            const int sw = (W * k.h + hmax - 1) / hmax, sh = (H * k.v + vmax - 1) / vmax;   // samples of the component
            k.cw = (sw + 7) / 8; k.ch = (sh + 7) / 8;
            k.bw = mcux * k.h; k.bh = mcuy * k.v;
            if ((size_t)k.bw * k.bh > ((size_t)1 << 26)) return fail("image too large");
            k.pix.assign((size_t)k.bw * 8 * k.bh * 8, 0);
            if (progressive) k.coef.assign((size_t)k.bw * k.bh * 64, 0);
        }
        return true;
    }

    // ---- inverse DCT ------------------------------------------------------------------------
    // LL&M integer transform, constants scaled by 2^12.  The even part works on (s0 +- s4) << 12,
    // the odd part is the usual four-rotation butterfly.  Pass 1 (columns) keeps two fractional
    // bits (>> 10 with rounding), pass 2 (rows) removes 17 bits, the rounding term also carrying
    // the +128 level shift; the result is clamped to a byte.
    static inline int fix(double x) { return (int)(x * 4096 + 0.5); }
    static inline uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

    struct Butterfly { int x0, x1, x2, x3, t0, t1, t2, t3; };
    static inline Butterfly butterfly(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7)
    {
        static const int c0541 = fix(0.5411961f), c1847 = fix(-1.847759065f), c0765 = fix(0.765366865f),
                         c1175 = fix(1.175875602f), c0298 = fix(0.298631336f), c2053 = fix(2.053119869f),
                         c3072 = fix(3.072711026f), c1501 = fix(1.501321110f), c0899 = fix(-0.899976223f),
                         c2562 = fix(-2.562915447f), c1961 = fix(-1.961570560f), c0390 = fix(-0.390180644f);
        Butterfly b;
        // even part
        int p1 = (s2 + s6) * c0541;
        const int e2 = p1 + s6 * c1847;
        const int e3 = p1 + s2 * c0765;
        const int e0 = (s0 + s4) * 4096;
        const int e1 = (s0 - s4) * 4096;
        b.x0 = e0 + e3; b.x3 = e0 - e3; b.x1 = e1 + e2; b.x2 = e1 - e2;
        // odd part
        int t0 = s7, t1 = s5, t2 = s3, t3 = s1;
        int p3 = t0 + t2, p4 = t1 + t3;
        p1 = t0 + t3;
        int p2 = t1 + t2;
        const int p5 = (p3 + p4) * c1175;
        t0 *= c0298; t1 *= c2053; t2 *= c3072; t3 *= c1501;
        p1 = p5 + p1 * c0899;
        p2 = p5 + p2 * c2562;
        p3 *= c1961;
        p4 *= c0390;
        b.t3 = t3 + p1 + p4;
        b.t2 = t2 + p2 + p3;
        b.t1 = t1 + p2 + p4;
        b.t0 = t0 + p1 + p3;
        return b;
    }

    static void idct8x8(uint8_t *out, int stride, const int16_t *blk)
    {
        int ws[64];
        for (int c = 0; c < 8; c++) {
            const int16_t *s = blk + c;
            int *w = ws + c;
            if (s[8] == 0 && s[16] == 0 && s[24] == 0 && s[32] == 0 && s[40] == 0 && s[48] == 0 && s[56] == 0) {
                const int dc = s[0] * 4;             // a DC-only column: the transform is a scale by 4
                w[0] = w[8] = w[16] = w[24] = w[32] = w[40] = w[48] = w[56] = dc;
            } else {
                Butterfly b = butterfly(s[0], s[8], s[16], s[24], s[32], s[40], s[48], s[56]);
                b.x0 += 512; b.x1 += 512; b.x2 += 512; b.x3 += 512;
                w[0] = (b.x0 + b.t3) >> 10;  w[56] = (b.x0 - b.t3) >> 10;
                w[8] = (b.x1 + b.t2) >> 10;  w[48] = (b.x1 - b.t2) >> 10;
                w[16] = (b.x2 + b.t1) >> 10; w[40] = (b.x2 - b.t1) >> 10;
                w[24] = (b.x3 + b.t0) >> 10; w[32] = (b.x3 - b.t0) >> 10;
            }
        }
        for (int r = 0; r < 8; r++) {
            const int *w = ws + 8 * r;
            uint8_t *o = out + (size_t)r * stride;
            Butterfly b = butterfly(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
            const int bias = 65536 + (128 << 17);    // rounding + level shift
            b.x0 += bias; b.x1 += bias; b.x2 += bias; b.x3 += bias;
            o[0] = clamp8((b.x0 + b.t3) >> 17); o[7] = clamp8((b.x0 - b.t3) >> 17);
            o[1] = clamp8((b.x1 + b.t2) >> 17); o[6] = clamp8((b.x1 - b.t2) >> 17);
            o[2] = clamp8((b.x2 + b.t1) >> 17); o[5] = clamp8((b.x2 - b.t1) >> 17);
            o[3] = clamp8((b.x3 + b.t0) >> 17); o[4] = clamp8((b.x3 - b.t0) >> 17);
        }
    }

    // ---- block decoders ---------------------------------------------------------------------
    // Sequential (T.81 F.2.2): dequantised while decoding, 16-bit coefficients.
    bool block_sequential(Component &k, int16_t *blk)
    {
        std::memset(blk, 0, 64 * sizeof(int16_t));
        const Huff &dc = hdc[k.td], &ac = hac[k.ta];
        const uint16_t *q = qt[k.tq];
        const int t = decode(dc);
        if (t < 0 || t > 15) return fail("bad DC code");
        k.dc_pred += receive_extend(t);
        blk[0] = (int16_t)(k.dc_pred * q[0]);
        for (int i = 1; i < 64;) {
            const int rs = decode(ac);
            if (rs < 0) return fail("bad AC code");
            const int r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (rs != 0xF0) break;          // EOB
                i += 16;
            } else {
                i += r;
                const int z = kZigzag[i++];
                blk[z] = (int16_t)(receive_extend(s) * q[z]);
            }
        }
        return true;
    }
    // Progressive DC scans (T.81 G.1.2.1)
    bool block_prog_dc(Component &k, int16_t *blk)
    {
        if (ah == 0) {
            const int t = decode(hdc[k.td]);
            if (t < 0 || t > 15) return fail("bad DC code");
            k.dc_pred += receive_extend(t);
            blk[0] = (int16_t)(k.dc_pred * (1 << al));
        } else if (getbit()) {
            blk[0] = (int16_t)(blk[0] + (1 << al));
        }
        return true;
    }
    // Progressive AC scans (T.81 G.1.2.2, G.1.2.3)
    bool block_prog_ac(Component &k, int16_t *blk)
    {
        const Huff &ac = hac[k.ta];
        if (ah == 0) {
            if (eobrun) { eobrun--; return true; }
            for (int i = ss; i <= se;) {
                const int rs = decode(ac);
                if (rs < 0) return fail("bad AC code");
                const int r = rs >> 4, s = rs & 15;
                if (s == 0) {
                    if (r < 15) {
                        eobrun = (1 << r) - 1;
                        if (r) eobrun += getbits(r);
                        break;
                    }
                    i += 16;
                } else {
                    i += r;
                    const int z = kZigzag[i++];
                    blk[z] = (int16_t)(receive_extend(s) * (1 << al));
                }
            }
            return true;
        }
        // refinement
        const int16_t bit = (int16_t)(1 << al);
        auto refine = [&](int16_t *p) {
            if (*p != 0 && getbit() && (*p & bit) == 0) *p = (int16_t)(*p > 0 ? *p + bit : *p - bit);
        };
        if (eobrun) {
            eobrun--;
            for (int i = ss; i <= se; i++) refine(&blk[kZigzag[i]]);
            return true;
        }
        int i = ss;
        do {
            const int rs = decode(ac);
            if (rs < 0) return fail("bad AC code");
            int r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (r < 15) {
                    eobrun = (1 << r) - 1;
                    if (r) eobrun += getbits(r);
                    r = 64;                      // refine the rest of the band, place nothing
                }
            } else {
                if (s != 1) return fail("bad refinement code");
                s = getbit() ? bit : -bit;
            }
            while (i <= se) {
                int16_t *p = &blk[kZigzag[i++]];
                if (*p != 0) refine(p);
                else {
                    if (r == 0) { *p = (int16_t)s; break; }
                    r--;
                }
            }
        } while (i <= se);
        return true;
    }

    bool restart_if_due(int &todo)
    {
        if (restart_interval == 0) return true;
        if (--todo > 0) return true;
        if (bitcnt < 24) fill();
        // T.81 E.2.4: an RSTm marker ends the interval
        if (!(marker >= 0xD0 && marker <= 0xD7)) { todo = 0x7fffffff; return true; }   // no restart marker: data ends
        reset_entropy();
        todo = restart_interval;
        return true;
    }

    bool read_scan(int len)
    {
        const int ns = u8();
        if (ns < 1 || ns > 4 || ns > ncomp || len != 4 + 2 * ns) return fail("bad SOS");
        int order[4];
        for (int s = 0; s < ns; s++) {
            const int id = u8(), tt = u8();
            int c = 0;
            while (c < ncomp && comp[c].id != id) c++;
            if (c == ncomp) return fail("bad SOS component");
            comp[c].td = tt >> 4; comp[c].ta = tt & 15;
            if (comp[c].td > 3 || comp[c].ta > 3) return fail("bad SOS table");
            order[s] = c;
        }
        ss = u8(); se = u8();
        const int a = u8();
        ah = a >> 4; al = a & 15;
        // a scan must not select a table no DHT segment defined (T.81 B.2.3): it would decode with empty code books
        for (int s = 0; s < ns; s++) {
            const Component &k = comp[order[s]];
            const bool need_dc = !progressive || (ss == 0 && ah == 0), need_ac = !progressive || se > 0;   // DC refinement scans read raw bits only
            if ((need_dc && !hdc[k.td].present) || (need_ac && !hac[k.ta].present)) return fail("scan uses an undefined Huffman table");
        }
        if (progressive) {
            if (ss > 63 || se > 63 || ss > se || ah > 13 || al > 13) return fail("bad progressive SOS");
            if (ss == 0 && se != 0) return fail("bad progressive SOS");
            if (ss != 0 && ns != 1) return fail("bad progressive SOS");
        } else {
            ss = 0; se = 63; ah = al = 0;
        }
        reset_entropy();
        int todo = restart_interval ? restart_interval : 0x7fffffff;
        int16_t blk[64];
        auto do_block = [&](Component &k, int bx, int by) -> bool {
            if (progressive) {
                int16_t *b = k.coef.data() + ((size_t)by * k.bw + bx) * 64;
                return ss == 0 ? block_prog_dc(k, b) : block_prog_ac(k, b);
            }
            if (!block_sequential(k, blk)) return false;
            idct8x8(k.pix.data() + ((size_t)by * 8 * k.bw + bx) * 8, k.bw * 8, blk);
            return true;
        };
        if (ns == 1) {
            // non-interleaved: the component's own blocks in raster order (T.81 A.2.2)
            Component &k = comp[order[0]];
            for (int by = 0; by < k.ch; by++)
                for (int bx = 0; bx < k.cw; bx++) {
                    if (!do_block(k, bx, by)) return false;
                    restart_if_due(todo);
                }
        } else {
            for (int my = 0; my < mcuy; my++)
                for (int mx = 0; mx < mcux; mx++) {
                    for (int s = 0; s < ns; s++) {
                        Component &k = comp[order[s]];
                        for (int y = 0; y < k.v; y++)
                            for (int x = 0; x < k.h; x++)
                                if (!do_block(k, mx * k.h + x, my * k.v + y)) return false;
                    }
                    restart_if_due(todo);
                }
        }
        // leave the entropy-coded segment: position on the next marker
        if (marker == 0) {
            while (pos + 1 < n && !(d[pos] == 0xFF && d[pos + 1] != 0 && d[pos + 1] != 0xFF && !(d[pos + 1] >= 0xD0 && d[pos + 1] <= 0xD7))) pos++;
        } else {
            pos -= 2;                           // the marker the bit reader consumed
            while (pos < n && !(d[pos] == 0xFF && pos + 1 < n && d[pos + 1] == marker)) pos++;
        }
        return true;
    }

    void finish_progressive()
    {
        int16_t blk[64];
        for (int c = 0; c < ncomp; c++) {
            Component &k = comp[c];
            const uint16_t *q = qt[k.tq];
            for (int by = 0; by < k.ch; by++)
                for (int bx = 0; bx < k.cw; bx++) {
                    const int16_t *src = k.coef.data() + ((size_t)by * k.bw + bx) * 64;
                    for (int i = 0; i < 64; i++) blk[i] = (int16_t)(src[i] * q[i]);     // 16-bit dequantisation
                    idct8x8(k.pix.data() + ((size_t)by * 8 * k.bw + bx) * 8, k.bw * 8, blk);
                }
        }
    }

    bool run(std::vector<uint8_t> &out, int *ow, int *oh, int *och)
    {
        if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return fail("no SOI");
        pos = 2;
        bool have_sof = false, have_scan = false;
        for (;;) {
            // next marker
            while (pos < n && d[pos] != 0xFF) pos++;
            while (pos < n && d[pos] == 0xFF) pos++;
            if (pos >= n) break;
            const int m = d[pos++];
            if (m == 0xD9) break;                                   // EOI
            if (m == 0x00 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
            if (pos + 2 > n) return fail("truncated");
            const int len = u16() - 2;
            if (len < 0 || pos + (size_t)len > n) return fail("truncated segment");
            const size_t next = pos + (size_t)len;
            if (m == 0xDB) { if (!read_dqt(len)) return false; }
            else if (m == 0xC4) { if (!read_dht(len)) return false; }
            else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
                if (have_sof) return fail("two SOF markers");
                progressive = (m == 0xC2);
                if (!read_sof(len)) return false;
                have_sof = true;
            } else if (m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
                return fail("unsupported coding process (lossless / hierarchical / arithmetic)");
            } else if (m == 0xDD) {
                if (len != 2) return fail("bad DRI");
                restart_interval = u16();
            } else if (m == 0xDA) {
                if (!have_sof) return fail("SOS before SOF");
                if (!read_scan(len)) return false;
                have_scan = true;
                continue;                                            // read_scan positioned us on the next marker
            }
            pos = next;
        }
        if (!have_sof || !have_scan) return fail("no image data");
        for (int c = 0; c < ncomp; c++)
            if (!qt_present[comp[c].tq]) return fail("missing quantisation table");
        if (progressive) finish_progressive();
        // gray output: the single component, or the luma plane of a colour file
        const Component &k = comp[0];
        out.resize((size_t)W * H);
        const int stride = k.bw * 8;
        for (int y = 0; y < H; y++) {
            const int sy = (int)((int64_t)y * k.v / vmax);
            const uint8_t *row = k.pix.data() + (size_t)sy * stride;
            if (k.h == hmax) std::memcpy(out.data() + (size_t)y * W, row, (size_t)W);
            else for (int x = 0; x < W; x++) out[(size_t)y * W + x] = row[(int64_t)x * k.h / hmax];
        }
        *ow = W; *oh = H; *och = ncomp;
        return true;
    }
};

}  // namespace

int jpeg_decode_gray(const uint8_t *data, size_t len, std::vector<uint8_t> &out, int *W, int *H, int *ch,
                     std::string &err)
{
    Decoder dec;
    dec.d = data; dec.n = len; dec.err = &err;
    std::memset(dec.qt, 0, sizeof(dec.qt));
    if (!dec.run(out, W, H, ch)) return DEFF2D_ERR_IO;
    return DEFF2D_OK;
}

}  // namespace deff2d
