// png_decode.cpp - PNG -> straight-alpha BGRA, rows top-down: what the reference's `new Bitmap(file)` + 32bppArgb conversion +
// ExtractBitmapBGRA (Engine/MeshLoaderOBJ.cs:463-502) hands to the scene for a PNG texture.
//
// The reference decodes every non-TGA image through System.Drawing (GDI+), a platform library that is not in the repository; PNG
// is lossless, so for the pixel formats whose GDI+ result is fully determined by the file this is an exact stand-in:
//   colour type 0 (grey, 1 / 2 / 4 / 8 bit), 2 (RGB 8), 3 (palette 1 / 2 / 4 / 8 bit, tRNS alpha), 4 (grey + alpha 8), 6 (RGBA 8),
//   tRNS colour keys for types 0 / 2, both interlace methods.  Non-premultiplied alpha throughout (Format32bppArgb is straight).
// Refused with InvalidDataException, not guessed: 16-bit samples (GDI+ loads them as 48 / 64bpp and converts through a
// platform-defined gamma step).  Ancillary colour-management chunks (gAMA, cHRM, iCCP, sRGB) are skipped, as GDI+ does for the
// usual sRGB / 1 / 2.2 files.  A damaged file (bad signature, CRC, zlib stream, sizes) raises ArgumentException, which is what
// `new Bitmap(file)` throws ("Parameter is not valid").
// Own inflate (RFC 1951: stored, fixed and dynamic Huffman blocks; canonical codes decoded bit by bit - textures are small) with
// the zlib wrapper's Adler-32 verified (RFC 1950).
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "engine.h"

namespace ILGPU_Raytracing {
namespace Engine {

namespace {

[[noreturn]] void bad(const std::string& file, const char* what) { throw ArgumentException("Parameter is not valid. (" + std::string(what) + ": " + file + ")"); }

struct Bits {
    const unsigned char* d; size_t n, pos = 0; uint32_t buf = 0; int cnt = 0; const std::string& file;
    int get(int need) {
        while (cnt < need) { if (pos >= n) bad(file, "zlib stream truncated"); buf |= (uint32_t)d[pos++] << cnt; cnt += 8; }
        const int v = (int)(buf & ((1u << need) - 1u));
        buf >>= need; cnt -= need;
        return v;
    }
    void align() { buf = 0; cnt = 0; }
};

// canonical Huffman code: count[len] codes of each length, symbols ordered by (length, symbol)
struct Huff {
    uint16_t count[16]; uint16_t symbol[288];
    bool build(const unsigned char* len, int n) {
        memset(count, 0, sizeof(count));
        for (int i = 0; i < n; i++) count[len[i]]++;
        if (count[0] == n) return true;   // no codes: legal only if never used
        int left = 1;
        for (int l = 1; l < 16; l++) { left <<= 1; left -= count[l]; if (left < 0) return false; }   // over-subscribed
        uint16_t offs[16]; offs[1] = 0;
        for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
        for (int i = 0; i < n; i++) if (len[i]) symbol[offs[len[i]]++] = (uint16_t)i;
        return true;
    }
    int decode(Bits& b) const {
        int code = 0, first = 0, index = 0;
        for (int l = 1; l < 16; l++) {
            code |= b.get(1);
            const int c = count[l];
            if (code - c < first) return symbol[index + (code - first)];
            index += c; first += c; first <<= 1; code <<= 1;
        }
        return -1;
    }
};

void inflate_codes(Bits& b, const Huff& lit, const Huff& dist, std::vector<unsigned char>& out, const std::string& file) {
    static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint16_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint16_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    for (;;) {
        int sym = lit.decode(b);
        if (sym < 0) bad(file, "bad Huffman code");
        if (sym < 256) { out.push_back((unsigned char)sym); continue; }
        if (sym == 256) return;
        sym -= 257;
        if (sym >= 29) bad(file, "bad length symbol");
        const int len = lbase[sym] + b.get(lext[sym]);
        const int ds = dist.decode(b);
        if (ds < 0 || ds >= 30) bad(file, "bad distance symbol");
        const size_t d = (size_t)dbase[ds] + (size_t)b.get(dext[ds]);
        if (d > out.size()) bad(file, "distance too far back");
        const size_t from = out.size() - d;
        for (int i = 0; i < len; i++) out.push_back(out[from + (size_t)i]);
    }
}

std::vector<unsigned char> zlib_inflate(const std::vector<unsigned char>& z, size_t expect, const std::string& file) {
    if (z.size() < 6) bad(file, "zlib stream too short");
    if ((z[0] & 0x0F) != 8 || ((z[0] << 8) | z[1]) % 31 != 0 || (z[1] & 0x20)) bad(file, "bad zlib header");
    Bits b{z.data() + 2, z.size() - 2, 0, 0, 0, file};
    std::vector<unsigned char> out; out.reserve(expect);
    int last;
    do {
        last = b.get(1);
        const int type = b.get(2);
        if (type == 0) {
            b.align();
            if (b.pos + 4 > b.n) bad(file, "stored block truncated");
            const unsigned len = b.d[b.pos] | (b.d[b.pos + 1] << 8), nlen = b.d[b.pos + 2] | (b.d[b.pos + 3] << 8);
            b.pos += 4;
            if ((len ^ 0xFFFFu) != nlen || b.pos + len > b.n) bad(file, "bad stored block");
            out.insert(out.end(), b.d + b.pos, b.d + b.pos + len); b.pos += len;
        } else if (type == 1) {
            unsigned char l[288]; Huff lit, dist;
            for (int i = 0; i < 144; i++) l[i] = 8; for (int i = 144; i < 256; i++) l[i] = 9; for (int i = 256; i < 280; i++) l[i] = 7; for (int i = 280; i < 288; i++) l[i] = 8;
            lit.build(l, 288);
            for (int i = 0; i < 30; i++) l[i] = 5;
            dist.build(l, 30);
            inflate_codes(b, lit, dist, out, file);
        } else if (type == 2) {
            static const unsigned char order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            const int nlen = b.get(5) + 257, ndist = b.get(5) + 1, ncode = b.get(4) + 4;
            if (nlen > 286 || ndist > 30) bad(file, "bad dynamic block header");
            unsigned char l[320]; memset(l, 0, sizeof(l));
            for (int i = 0; i < ncode; i++) l[order[i]] = (unsigned char)b.get(3);
            Huff cl;
            if (!cl.build(l, 19)) bad(file, "bad code-length code");
            unsigned char lens[320]; int idx = 0;
            while (idx < nlen + ndist) {
                const int sym = cl.decode(b);
                if (sym < 0) bad(file, "bad code-length symbol");
                if (sym < 16) { lens[idx++] = (unsigned char)sym; continue; }
                int rep, val = 0;
                if (sym == 16) { if (idx == 0) bad(file, "repeat without a previous length"); val = lens[idx - 1]; rep = 3 + b.get(2); }
                else if (sym == 17) rep = 3 + b.get(3);
                else rep = 11 + b.get(7);
                if (idx + rep > nlen + ndist) bad(file, "too many code lengths");
                while (rep--) lens[idx++] = (unsigned char)val;
            }
            if (lens[256] == 0) bad(file, "no end-of-block code");
            Huff lit, dist;
            if (!lit.build(lens, nlen) || !dist.build(lens + nlen, ndist)) bad(file, "over-subscribed Huffman code");
            inflate_codes(b, lit, dist, out, file);
        } else bad(file, "bad block type");
    } while (!last);
    b.align();
    if (b.pos + 4 > b.n) bad(file, "Adler-32 missing");
    uint32_t a = 1, s = 0;
    for (unsigned char c : out) { a = (a + c) % 65521u; s = (s + a) % 65521u; }
    const uint32_t want = ((uint32_t)b.d[b.pos] << 24) | ((uint32_t)b.d[b.pos + 1] << 16) | ((uint32_t)b.d[b.pos + 2] << 8) | b.d[b.pos + 3];
    if (((s << 16) | a) != want) bad(file, "Adler-32 mismatch");
    return out;
}

uint32_t crc32(const unsigned char* p, size_t n) {
    static uint32_t table[256]; static bool made = false;
    if (!made) { for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; } made = true; }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}
uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
int paeth(int a, int b, int c) { const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p; return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }

}   // namespace

TextureSrc load_png_bgra(const std::string& file, const std::vector<unsigned char>& bytes) {
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (bytes.size() < 8 || memcmp(bytes.data(), sig, 8) != 0) bad(file, "not a PNG file");
    size_t p = 8;
    bool haveHdr = false, ended = false;
    uint32_t w = 0, h = 0; int depth = 0, ctype = 0, interlace = 0;
    std::vector<unsigned char> plte, trns, idat;
    while (!ended) {
        if (p + 12 > bytes.size()) bad(file, "chunk truncated");
        const uint32_t len = be32(&bytes[p]);
        if ((size_t)len > bytes.size() - p - 12) bad(file, "chunk truncated");
        const unsigned char* type = &bytes[p + 4]; const unsigned char* data = &bytes[p + 8];
        if (crc32(type, (size_t)len + 4) != be32(data + len)) bad(file, "chunk CRC mismatch");
        if (memcmp(type, "IHDR", 4) == 0) {
            if (len != 13 || haveHdr) bad(file, "bad IHDR");
            w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
            if (w == 0 || h == 0 || w > 32768u || h > 32768u || data[10] != 0 || data[11] != 0 || interlace > 1) bad(file, "bad IHDR");
            haveHdr = true;
        } else if (!haveHdr) bad(file, "IHDR is not the first chunk");
        else if (memcmp(type, "PLTE", 4) == 0) { if (len % 3 != 0 || len > 768) bad(file, "bad PLTE"); plte.assign(data, data + len); }
        else if (memcmp(type, "tRNS", 4) == 0) trns.assign(data, data + len);
        else if (memcmp(type, "IDAT", 4) == 0) idat.insert(idat.end(), data, data + len);
        else if (memcmp(type, "IEND", 4) == 0) ended = true;
        else if (!(type[0] & 0x20)) bad(file, "unknown critical chunk");   // ancillary chunks (gAMA, sRGB, tEXt, pHYs ...) are skipped
        p += (size_t)len + 12;
    }
    int channels;
    switch (ctype) {
        case 0: channels = 1; if (depth != 1 && depth != 2 && depth != 4 && depth != 8 && depth != 16) bad(file, "bad bit depth"); break;
        case 2: channels = 3; if (depth != 8 && depth != 16) bad(file, "bad bit depth"); break;
        case 3: channels = 1; if (depth != 1 && depth != 2 && depth != 4 && depth != 8) bad(file, "bad bit depth"); if (plte.empty()) bad(file, "palette missing"); break;
        case 4: channels = 2; if (depth != 8 && depth != 16) bad(file, "bad bit depth"); break;
        case 6: channels = 4; if (depth != 8 && depth != 16) bad(file, "bad bit depth"); break;
        default: bad(file, "bad colour type");
    }
    if (depth == 16) throw InvalidDataException("16-bit PNG '" + file + "': System.Drawing converts 48 / 64bpp images through a platform-defined gamma step; use an 8-bit PNG, TGA or BMP");
    const int bitsPerPixel = channels * depth, bpp = bitsPerPixel >= 8 ? bitsPerPixel / 8 : 1;
    // passes: the whole image, or the seven Adam7 sub-images
    static const int px0[7] = {0, 4, 0, 2, 0, 1, 0}, py0[7] = {0, 0, 4, 0, 2, 0, 1}, pdx[7] = {8, 8, 4, 4, 2, 2, 1}, pdy[7] = {8, 8, 8, 4, 4, 2, 2};
    const int nPass = interlace ? 7 : 1;
    size_t expect = 0;
    for (int k = 0; k < nPass; k++) {
        const uint32_t pw = interlace ? (w - px0[k] + pdx[k] - 1) / pdx[k] : w, ph = interlace ? (h - py0[k] + pdy[k] - 1) / pdy[k] : h;
        if (interlace && ((int)w <= px0[k] || (int)h <= py0[k])) continue;
        expect += (size_t)ph * (1 + ((size_t)pw * bitsPerPixel + 7) / 8);
    }
    std::vector<unsigned char> raw = zlib_inflate(idat, expect, file);
    if (raw.size() < expect) bad(file, "image data too short");

    TextureSrc tex; tex.Path = file; tex.Width = (int)w; tex.Height = (int)h; tex.BGRA.assign((size_t)w * h * 4, 0);
    auto put = [&](uint32_t x, uint32_t y, const unsigned char* s /* `channels` samples, 8 bit each (palette: the index) */) {
        unsigned char* q = &tex.BGRA[((size_t)y * w + x) * 4];
        switch (ctype) {
            case 0: q[0] = q[1] = q[2] = s[0]; q[3] = 255; break;
            case 2: q[0] = s[2]; q[1] = s[1]; q[2] = s[0]; q[3] = 255; break;
            case 3: { const size_t i = s[0]; if (i * 3 + 2 >= plte.size()) bad(file, "palette index out of range");
                      q[0] = plte[i * 3 + 2]; q[1] = plte[i * 3 + 1]; q[2] = plte[i * 3]; q[3] = i < trns.size() ? trns[i] : 255; break; }
            case 4: q[0] = q[1] = q[2] = s[0]; q[3] = s[1]; break;
            default: q[0] = s[2]; q[1] = s[1]; q[2] = s[0]; q[3] = s[3]; break;
        }
    };
    // tRNS colour keys (types 0 and 2): the one colour that is fully transparent, compared on the file's sample values
    int keyGrey = -1, keyR = -1, keyG = -1, keyB = -1;
    if (ctype == 0 && trns.size() >= 2) keyGrey = (trns[0] << 8) | trns[1];
    if (ctype == 2 && trns.size() >= 6) { keyR = (trns[0] << 8) | trns[1]; keyG = (trns[2] << 8) | trns[3]; keyB = (trns[4] << 8) | trns[5]; }
    size_t at = 0;
    std::vector<unsigned char> prev, cur;
    for (int k = 0; k < nPass; k++) {
        if (interlace && ((int)w <= px0[k] || (int)h <= py0[k])) continue;
        const uint32_t pw = interlace ? (w - px0[k] + pdx[k] - 1) / pdx[k] : w, ph = interlace ? (h - py0[k] + pdy[k] - 1) / pdy[k] : h;
        const size_t rowBytes = ((size_t)pw * bitsPerPixel + 7) / 8;
        prev.assign(rowBytes, 0); cur.assign(rowBytes, 0);
        for (uint32_t ry = 0; ry < ph; ry++) {
            const int filter = raw[at++];
            const unsigned char* src = &raw[at]; at += rowBytes;
            if (filter > 4) bad(file, "bad filter type");
            for (size_t i = 0; i < rowBytes; i++) {
                const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
                int v = src[i];
                if (filter == 1) v += a; else if (filter == 2) v += b; else if (filter == 3) v += (a + b) >> 1; else if (filter == 4) v += paeth(a, b, c);
                cur[i] = (unsigned char)v;
            }
            const uint32_t y = interlace ? py0[k] + ry * pdy[k] : ry;
            for (uint32_t rx = 0; rx < pw; rx++) {
                const uint32_t x = interlace ? px0[k] + rx * pdx[k] : rx;
                unsigned char s[4] = {0, 0, 0, 0};
                if (depth == 8) { for (int ch = 0; ch < channels; ch++) s[ch] = cur[(size_t)rx * channels + ch]; }
                else {   // 1 / 2 / 4 bit samples, most significant bits first (grey or palette: one channel)
                    const size_t bit = (size_t)rx * depth;
                    const int v = (cur[bit >> 3] >> (8 - depth - (int)(bit & 7))) & ((1 << depth) - 1);
                    s[0] = (unsigned char)(ctype == 3 ? v : v * 255 / ((1 << depth) - 1));
                    if (ctype == 0 && keyGrey == v) { put(x, y, s); tex.BGRA[((size_t)y * w + x) * 4 + 3] = 0; continue; }
                }
                put(x, y, s);
                if (depth == 8 && ctype == 0 && keyGrey == s[0]) tex.BGRA[((size_t)y * w + x) * 4 + 3] = 0;
                if (ctype == 2 && keyR == s[0] && keyG == s[1] && keyB == s[2]) tex.BGRA[((size_t)y * w + x) * 4 + 3] = 0;
            }
            prev.swap(cur);
        }
    }
    return tex;
}

}   // namespace Engine
}   // namespace ILGPU_Raytracing
