// bmfr_io — dataset ingestion for the driver program (include/bmfr_io.h).
//
// The reference reads its inputs with OpenImageIO (read_image_file, bmfr.cpp:145-165), compiles the
// dataset's camera_matrices.h into the binary (bmfr.cpp:46-47) and writes PNGs with OpenImageIO
// (bmfr.cpp:520-539).  None of that is available here, so this file holds a small OpenEXR scanline
// reader, a tolerant parser for the header's initialisers and a PNG encoder on top of zlib.
#include "../../include/bmfr_io.h"

#include <ctype.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <string>
#include <vector>

namespace {

thread_local char g_error[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

bool read_file(const char* path, std::vector<unsigned char>& out) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n < 0) {
        fclose(f);
        return false;
    }
    out.resize((size_t)n);
    const size_t got = n ? fread(out.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)n;
}

// ---- OpenEXR ---------------------------------------------------------------------------------
enum { PIXEL_UINT = 0, PIXEL_HALF = 1, PIXEL_FLOAT = 2 };
enum { COMP_NONE = 0, COMP_RLE = 1, COMP_ZIPS = 2, COMP_ZIP = 3, COMP_PIZ = 4 };

struct Channel {
    std::string name;
    int type = 0, xs = 1, ys = 1;
};
struct ExrHeader {
    std::vector<Channel> channels;  // in file order = alphabetical
    int compression = -1, line_order = 0;
    int xmin = 0, ymin = 0, xmax = -1, ymax = -1;
    size_t table = 0;  // file offset of the chunk offset table
    // (64-bit: the four corners are arbitrary 32-bit values in a damaged file)
    long long width() const { return (long long)xmax - xmin + 1; }
    long long height() const { return (long long)ymax - ymin + 1; }
};

struct Cursor {
    const unsigned char* p;
    size_t n, at = 0;
    bool has(size_t k) const { return at + k <= n; }
    bool str(std::string& s, size_t max_len = 255) {
        s.clear();
        while (at < n && p[at] != 0) {
            if (s.size() >= max_len) return false;
            s.push_back((char)p[at++]);
        }
        if (at >= n) return false;
        ++at;
        return true;
    }
    bool i32(int32_t& v) {
        if (!has(4)) return false;
        uint32_t u = (uint32_t)p[at] | ((uint32_t)p[at + 1] << 8) | ((uint32_t)p[at + 2] << 16) | ((uint32_t)p[at + 3] << 24);
        v = (int32_t)u;
        at += 4;
        return true;
    }
};

int parse_header(const std::vector<unsigned char>& file, const char* path, ExrHeader& h) {
    Cursor c{file.data(), file.size()};
    int32_t magic = 0, version = 0;
    if (!c.i32(magic) || !c.i32(version) || magic != 20000630)
        return fail(BMFR_IO_ERR_FORMAT, "%s: not an OpenEXR file", path);
    if ((version & 0xff) != 2) return fail(BMFR_IO_ERR_UNSUPPORTED, "%s: OpenEXR version %d", path, version & 0xff);
    if (version & 0x200) return fail(BMFR_IO_ERR_UNSUPPORTED, "%s: tiled OpenEXR files are not covered", path);
    if (version & (0x800 | 0x1000)) return fail(BMFR_IO_ERR_UNSUPPORTED, "%s: deep / multi-part OpenEXR files are not covered", path);
    bool have_channels = false, have_window = false;
    for (;;) {
        std::string name, type;
        if (!c.str(name)) return fail(BMFR_IO_ERR_FORMAT, "%s: truncated header", path);
        if (name.empty()) break;
        int32_t size = 0;
        if (!c.str(type) || !c.i32(size) || size < 0 || !c.has((size_t)size))
            return fail(BMFR_IO_ERR_FORMAT, "%s: truncated attribute '%s'", path, name.c_str());
        Cursor a{file.data() + c.at, (size_t)size};
        c.at += (size_t)size;
        if (name == "channels") {
            for (;;) {
                Channel ch;
                if (!a.str(ch.name)) return fail(BMFR_IO_ERR_FORMAT, "%s: truncated channel list", path);
                if (ch.name.empty()) break;
                int32_t t = 0, xs = 0, ys = 0;
                if (!a.i32(t) || !a.has(4)) return fail(BMFR_IO_ERR_FORMAT, "%s: truncated channel list", path);
                a.at += 4;  // pLinear + 3 reserved bytes
                if (!a.i32(xs) || !a.i32(ys)) return fail(BMFR_IO_ERR_FORMAT, "%s: truncated channel list", path);
                ch.type = t; ch.xs = xs; ch.ys = ys;
                h.channels.push_back(ch);
            }
            have_channels = true;
        } else if (name == "compression") {
            if (size < 1) return fail(BMFR_IO_ERR_FORMAT, "%s: bad compression attribute", path);
            h.compression = a.p[0];
        } else if (name == "dataWindow") {
            int32_t v[4];
            for (int i = 0; i < 4; ++i)
                if (!a.i32(v[i])) return fail(BMFR_IO_ERR_FORMAT, "%s: bad dataWindow", path);
            h.xmin = v[0]; h.ymin = v[1]; h.xmax = v[2]; h.ymax = v[3];
            have_window = true;
        } else if (name == "lineOrder") {
            if (size < 1) return fail(BMFR_IO_ERR_FORMAT, "%s: bad lineOrder attribute", path);
            h.line_order = a.p[0];
        }
    }
    if (!have_channels || !have_window || h.compression < 0)
        return fail(BMFR_IO_ERR_FORMAT, "%s: header lacks channels / dataWindow / compression", path);
    if (h.width() <= 0 || h.height() <= 0 || h.width() > (1 << 20) || h.height() > (1 << 20))
        return fail(BMFR_IO_ERR_FORMAT, "%s: implausible data window", path);
    h.table = c.at;
    return BMFR_IO_OK;
}

float half_to_float(uint16_t v) {
    const uint32_t sign = (uint32_t)(v & 0x8000u) << 16;
    uint32_t exp = (v >> 10) & 0x1fu, man = v & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {  // subnormal: normalise
            int e = -1;
            do {
                ++e;
                man <<= 1;
            } while ((man & 0x400u) == 0);
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

// The byte transform OpenEXR applies before zlib / RLE: undo the delta predictor, then interleave
// the two halves back.
void unpredict_and_interleave(std::vector<unsigned char>& buf, std::vector<unsigned char>& tmp) {
    const size_t n = buf.size();
    if (n == 0) return;
    for (size_t i = 1; i < n; ++i) buf[i] = (unsigned char)(buf[i - 1] + buf[i] - 128);
    tmp.resize(n);
    const size_t half = (n + 1) / 2;
    for (size_t i = 0; i < n; ++i) tmp[i] = (i & 1) ? buf[half + i / 2] : buf[i / 2];
    buf.swap(tmp);
}

bool rle_decode(const unsigned char* src, size_t n, std::vector<unsigned char>& out, size_t expect) {
    out.clear();
    size_t i = 0;
    while (i < n) {
        const int count = (signed char)src[i++];
        if (count < 0) {
            const size_t k = (size_t)(-count);
            if (i + k > n || out.size() + k > expect) return false;
            out.insert(out.end(), src + i, src + i + k);
            i += k;
        } else {
            if (i >= n || out.size() + (size_t)count + 1 > expect) return false;
            out.insert(out.end(), (size_t)count + 1, src[i++]);
        }
    }
    return out.size() == expect;
}

// ---- PIZ: 16-bit value range compaction (bitmap + LUT), 2-D Haar-like wavelet, Huffman ----------
// Decoder for OpenEXR's PIZ chunks, written from the format: the chunk holds the range of used 16-bit
// values as a bitmap, then a canonical-Huffman stream (6-bit packed code lengths with zero runs, one
// run-length symbol) of the wavelet-transformed, channel-planar 16-bit planes (a FLOAT channel = two
// interleaved planes).
struct BitReader {
    const unsigned char* p;
    const unsigned char* end;
    uint64_t c = 0;
    int lc = 0;
    bool ok = true;
    uint32_t get(int n) {  // MSB first
        while (lc < n) {
            if (p >= end) {
                ok = false;
                return 0;
            }
            c = (c << 8) | *p++;
            lc += 8;
        }
        lc -= n;
        return (uint32_t)((c >> lc) & ((1ull << n) - 1));
    }
};

constexpr int kHufEncSize = (1 << 16) + 1;  // symbols 0..65535 and the run-length symbol

bool huf_uncompress(const unsigned char* in, size_t n_in, uint16_t* out, size_t n_out) {
    if (n_in < 20) return n_out == 0;
    auto u32 = [&](size_t at) { return (uint32_t)in[at] | ((uint32_t)in[at + 1] << 8) | ((uint32_t)in[at + 2] << 16) | ((uint32_t)in[at + 3] << 24); };
    const uint32_t im = u32(0), iM = u32(4), n_bits = u32(12);
    if (im >= (uint32_t)kHufEncSize || iM >= (uint32_t)kHufEncSize || im > iM) return false;
    // code lengths im..iM: 6 bits each; 59..62 = run of 2..5 zeros, 63 = run of 6 + next 8 bits zeros
    std::vector<unsigned char> len((size_t)kHufEncSize, 0);
    BitReader br{in + 20, in + n_in};
    for (uint32_t i = im; i <= iM; ++i) {
        const uint32_t l = br.get(6);
        if (!br.ok) return false;
        if (l == 63) {
            uint32_t run = br.get(8) + 6;
            if (!br.ok || i + run > iM + 1) return false;
            i += run - 1;
        } else if (l >= 59) {
            const uint32_t run = l - 59 + 2;
            if (i + run > iM + 1) return false;
            i += run - 1;
        } else {
            len[i] = (unsigned char)l;
        }
    }
    // canonical codes: within a length codes grow with the symbol, longer lengths take the numerically
    // smaller prefixes (first[l] = (first[l+1] + count[l+1]) >> 1)
    uint64_t count[59] = {0}, first[60] = {0};
    for (uint32_t i = im; i <= iM; ++i) count[len[i]]++;
    {
        uint64_t c = 0;
        for (int l = 58; l >= 1; --l) {
            const uint64_t nc = (c + count[l]) >> 1;
            first[l] = c;
            c = nc;
        }
    }
    size_t offset[60] = {0};
    for (int l = 1; l < 59; ++l) offset[l + 1] = offset[l] + (size_t)count[l];
    std::vector<uint32_t> syms(offset[59]);
    {
        size_t fill[60];
        memcpy(fill, offset, sizeof(fill));
        for (uint32_t i = im; i <= iM; ++i)
            if (len[i]) syms[fill[len[i]]++] = i;
    }
    // the data starts at the next byte boundary after the table
    const unsigned char* data = br.p;  // bytes consumed so far (partial byte bits are discarded)
    BitReader dr{data, in + n_in};
    uint64_t left = n_bits;
    size_t o = 0;
    while (left > 0) {
        uint64_t code = 0;
        int l = 0;
        uint32_t sym = 0;
        for (;;) {
            if (left == 0 || l >= 58) return false;
            code = (code << 1) | dr.get(1);
            if (!dr.ok) return false;
            --left;
            ++l;
            if (count[l] && code >= first[l] && code - first[l] < count[l]) {
                sym = syms[offset[l] + (size_t)(code - first[l])];
                break;
            }
        }
        if (sym == iM) {  // run-length symbol: repeat the previous value
            if (left < 8 || o == 0) return false;
            const uint32_t run = dr.get(8);
            if (!dr.ok) return false;
            left -= 8;
            if (o + run > n_out) return false;
            for (uint32_t r = 0; r < run; ++r, ++o) out[o] = out[o - 1];
        } else {
            if (o >= n_out) return false;
            out[o++] = (uint16_t)sym;
        }
    }
    return o == n_out;
}

inline void wdec14(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
    const int ls = (int16_t)l, hs = (int16_t)h;
    const int ai = ls + (hs & 1) + (hs >> 1);
    a = (uint16_t)(int16_t)ai;
    b = (uint16_t)(int16_t)(ai - hs);
}
inline void wdec16(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
    const int m = l, d = h;
    const int bb = (m - (d >> 1)) & 0xffff;
    const int aa = (d + bb - (1 << 15)) & 0xffff;
    b = (uint16_t)bb;
    a = (uint16_t)aa;
}
// in: nx x ny values with strides ox, oy (in uint16 units); mx = largest value after the LUT
void wav2_decode(uint16_t* in, int nx, int ox, int ny, int oy, uint16_t mx) {
    const bool w14 = mx < (1 << 14);
    const int n = nx > ny ? ny : nx;
    int p = 1, p2;
    while (p <= n) p <<= 1;
    p >>= 1;
    p2 = p;
    p >>= 1;
    auto dec = [&](uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) {
        if (w14) wdec14(l, h, a, b);
        else wdec16(l, h, a, b);
    };
    while (p >= 1) {
        uint16_t* py = in;
        uint16_t* ey = in + (ptrdiff_t)oy * (ny - p2);
        const ptrdiff_t oy1 = (ptrdiff_t)oy * p, oy2 = (ptrdiff_t)oy * p2, ox1 = (ptrdiff_t)ox * p, ox2 = (ptrdiff_t)ox * p2;
        uint16_t i00, i01, i10, i11;
        for (; py <= ey; py += oy2) {
            uint16_t* px = py;
            uint16_t* ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t* p01 = px + ox1;
                uint16_t* p10 = px + oy1;
                uint16_t* p11 = p10 + ox1;
                dec(*px, *p10, i00, i10);
                dec(*p01, *p11, i01, i11);
                dec(i00, i01, *px, *p01);
                dec(i10, i11, *p10, *p11);
            }
            if (nx & p) {
                uint16_t* p10 = px + oy1;
                dec(*px, *p10, i00, *p10);
                *px = i00;
            }
        }
        if (ny & p) {
            uint16_t* px = py;
            uint16_t* ex = py + (ptrdiff_t)ox * (nx - p2);
            for (; px <= ex; px += ox2) {
                uint16_t* p01 = px + ox1;
                dec(*px, *p01, i00, *p01);
                *px = i00;
            }
        }
        p2 = p;
        p >>= 1;
    }
}

// One PIZ chunk -> `lines` scanlines in the raw layout (per scanline: channel after channel).
bool piz_decode(const unsigned char* src, size_t n, const std::vector<Channel>& ch, int width, int lines, std::vector<unsigned char>& raw) {
    if (n < 4) return false;
    const uint32_t min_nz = src[0] | (src[1] << 8), max_nz = src[2] | (src[3] << 8);
    size_t at = 4;
    std::vector<unsigned char> bitmap(8192, 0);
    if (max_nz >= 8192) return false;
    if (min_nz <= max_nz) {
        const size_t nb = max_nz - min_nz + 1;
        if (at + nb > n) return false;
        memcpy(bitmap.data() + min_nz, src + at, nb);
        at += nb;
    }
    std::vector<uint16_t> lut(65536, 0);
    uint32_t k = 0;
    for (uint32_t i = 0; i < 65536; ++i)
        if (i == 0 || (bitmap[i >> 3] & (1 << (i & 7)))) lut[k++] = (uint16_t)i;
    const uint16_t max_value = (uint16_t)(k - 1);
    if (at + 4 > n) return false;
    const uint32_t hlen = (uint32_t)src[at] | ((uint32_t)src[at + 1] << 8) | ((uint32_t)src[at + 2] << 16) | ((uint32_t)src[at + 3] << 24);
    at += 4;
    if (hlen > n - at) return false;
    size_t total = 0;  // uint16 values of the chunk
    for (const Channel& c : ch) total += (size_t)width * lines * (c.type == PIXEL_HALF ? 1 : 2);
    std::vector<uint16_t> planar(total);
    if (!huf_uncompress(src + at, hlen, planar.data(), total)) return false;
    size_t start = 0;
    std::vector<size_t> starts;
    for (const Channel& c : ch) {
        const int size = c.type == PIXEL_HALF ? 1 : 2;
        starts.push_back(start);
        for (int j = 0; j < size; ++j) wav2_decode(planar.data() + start + j, width, size, lines, width * size, max_value);
        start += (size_t)width * lines * size;
    }
    for (uint16_t& v : planar) v = lut[v];
    raw.resize(total * 2);
    unsigned char* out = raw.data();
    for (int y = 0; y < lines; ++y)
        for (size_t i = 0; i < ch.size(); ++i) {
            const size_t nvals = (size_t)width * (ch[i].type == PIXEL_HALF ? 1 : 2);
            const uint16_t* row = planar.data() + starts[i] + (size_t)y * nvals;
            for (size_t v = 0; v < nvals; ++v) {  // little-endian, whatever the host is
                *out++ = (unsigned char)(row[v] & 0xff);
                *out++ = (unsigned char)(row[v] >> 8);
            }
        }
    return true;
}

// Which file channel feeds output component 0, 1, 2.
bool pick_rgb(const std::vector<Channel>& ch, int order[3]) {
    auto base = [](const std::string& s) {
        const size_t dot = s.rfind('.');
        return dot == std::string::npos ? s : s.substr(dot + 1);
    };
    const char* sets[2][3] = {{"R", "G", "B"}, {"X", "Y", "Z"}};
    for (auto& set : sets) {
        int found[3] = {-1, -1, -1};
        for (int k = 0; k < 3; ++k)
            for (int i = 0; i < (int)ch.size(); ++i) {
                std::string b = base(ch[i].name);
                for (auto& chr : b) chr = (char)toupper((unsigned char)chr);
                if (b == set[k]) found[k] = i;
            }
        if (found[0] >= 0 && found[1] >= 0 && found[2] >= 0) {
            order[0] = found[0]; order[1] = found[1]; order[2] = found[2];
            return true;
        }
    }
    order[0] = 0; order[1] = 1; order[2] = 2;
    return ch.size() >= 3;
}

// ---- camera_matrices.h -----------------------------------------------------------------------
std::string strip_comments(const std::string& s) {
    std::string out;
    out.reserve(s.size());
    for (size_t i = 0; i < s.size();) {
        if (s.compare(i, 2, "//") == 0) {
            while (i < s.size() && s[i] != '\n') ++i;
        } else if (s.compare(i, 2, "/*") == 0) {
            const size_t e = s.find("*/", i + 2);
            i = e == std::string::npos ? s.size() : e + 2;
            out.push_back(' ');
        } else {
            out.push_back(s[i++]);
        }
    }
    return out;
}

bool ident_char(char c) { return isalnum((unsigned char)c) || c == '_'; }

// The numeric literals of `name ... = <initialiser> ;`.  Returns false when the name is not defined.
bool initialiser_values(const std::string& text, const char* name, std::vector<float>& values) {
    const size_t len = strlen(name);
    for (size_t at = text.find(name); at != std::string::npos; at = text.find(name, at + 1)) {
        if (at > 0 && ident_char(text[at - 1])) continue;
        if (at + len < text.size() && ident_char(text[at + len])) continue;
        const size_t eq = text.find_first_of("=;", at + len);
        if (eq == std::string::npos || text[eq] != '=') continue;  // a declaration or a use, not a definition
        const size_t end = text.find(';', eq);
        const std::string init = text.substr(eq + 1, (end == std::string::npos ? text.size() : end) - eq - 1);
        values.clear();
        const char* p = init.c_str();
        while (*p) {
            if (isdigit((unsigned char)*p) || ((*p == '-' || *p == '+' || *p == '.') && (isdigit((unsigned char)p[1]) || p[1] == '.'))) {
                char* e = nullptr;
                const double v = strtod(p, &e);
                if (e == p) {
                    ++p;
                    continue;
                }
                values.push_back((float)v);
                p = e;
                while (*p == 'f' || *p == 'F' || *p == 'l' || *p == 'L') ++p;
            } else if (ident_char(*p)) {
                while (ident_char(*p)) ++p;  // an identifier (a macro, a cast): skip it whole, digits included
            } else {
                ++p;
            }
        }
        return true;
    }
    return false;
}

// ---- PNG -------------------------------------------------------------------------------------
void put_be32(std::vector<unsigned char>& v, uint32_t x) {
    v.push_back((unsigned char)(x >> 24)); v.push_back((unsigned char)(x >> 16));
    v.push_back((unsigned char)(x >> 8)); v.push_back((unsigned char)x);
}
void put_chunk(std::vector<unsigned char>& png, const char type[4], const unsigned char* data, size_t n) {
    put_be32(png, (uint32_t)n);
    const size_t start = png.size();
    png.insert(png.end(), type, type + 4);
    if (n) png.insert(png.end(), data, data + n);
    put_be32(png, (uint32_t)crc32(0L, png.data() + start, (uInt)(png.size() - start)));
}

}  // namespace

extern "C" {

const char* bmfr_io_last_error(void) { return g_error; }

int bmfr_io_exr_info(const char* path, int* width, int* height, int* channels) {
    if (!path || !width || !height || !channels) return fail(BMFR_IO_ERR_ARGUMENT, "bmfr_io_exr_info: null argument");
    std::vector<unsigned char> file;
    if (!read_file(path, file)) return fail(BMFR_IO_ERR_OPEN, "%s: cannot open", path);
    ExrHeader h;
    const int st = parse_header(file, path, h);
    if (st != BMFR_IO_OK) return st;
    *width = (int)h.width();
    *height = (int)h.height();
    *channels = (int)h.channels.size();
    return BMFR_IO_OK;
}

int bmfr_io_read_exr_rgb(const char* path, int width, int height, float* rgb) {
    if (!path || !rgb || width <= 0 || height <= 0) return fail(BMFR_IO_ERR_ARGUMENT, "bmfr_io_read_exr_rgb: bad argument");
    std::vector<unsigned char> file;
    if (!read_file(path, file)) return fail(BMFR_IO_ERR_OPEN, "%s: cannot open", path);
    ExrHeader h;
    int st = parse_header(file, path, h);
    if (st != BMFR_IO_OK) return st;
    if (h.width() != width || h.height() != height || h.channels.size() != 3)  // bmfr.cpp:150-155
        return fail(BMFR_IO_ERR_MISMATCH, "%s: %d x %d with %d channels, expected %d x %d with 3 (wrong type)", path, (int)h.width(),
                    (int)h.height(), (int)h.channels.size(), width, height);
    if (h.compression > COMP_PIZ)
        return fail(BMFR_IO_ERR_UNSUPPORTED, "%s: compression %d is not covered (NONE, RLE, ZIPS, ZIP, PIZ are)", path, h.compression);
    size_t line_bytes = 0;
    std::vector<size_t> ch_offset(h.channels.size());
    for (size_t i = 0; i < h.channels.size(); ++i) {
        const Channel& c = h.channels[i];
        if (c.xs != 1 || c.ys != 1) return fail(BMFR_IO_ERR_UNSUPPORTED, "%s: subsampled channel '%s'", path, c.name.c_str());
        if (c.type != PIXEL_HALF && c.type != PIXEL_FLOAT)
            return fail(BMFR_IO_ERR_UNSUPPORTED, "%s: channel '%s' is neither HALF nor FLOAT", path, c.name.c_str());
        ch_offset[i] = line_bytes;
        line_bytes += (size_t)width * (c.type == PIXEL_HALF ? 2 : 4);
    }
    int order[3];
    pick_rgb(h.channels, order);

    const int lines_per_chunk = h.compression == COMP_ZIP ? 16 : h.compression == COMP_PIZ ? 32 : 1;
    const int chunks = (height + lines_per_chunk - 1) / lines_per_chunk;
    if (h.table + (size_t)chunks * 8 > file.size()) return fail(BMFR_IO_ERR_FORMAT, "%s: truncated offset table", path);
    std::vector<unsigned char> raw, tmp;
    std::vector<char> seen((size_t)height, 0);
    for (int k = 0; k < chunks; ++k) {
        uint64_t off = 0;
        for (int b = 0; b < 8; ++b) off |= (uint64_t)file[h.table + (size_t)k * 8 + b] << (8 * b);
        if (file.size() < 8 || off > file.size() - 8) return fail(BMFR_IO_ERR_FORMAT, "%s: chunk %d lies outside the file", path, k);
        Cursor c{file.data(), file.size(), (size_t)off};
        int32_t y = 0, size = 0;
        c.i32(y);
        c.i32(size);
        if (size < 0 || !c.has((size_t)size)) return fail(BMFR_IO_ERR_FORMAT, "%s: chunk %d is truncated", path, k);
        const long long row = (long long)y - h.ymin;
        if (row < 0 || row >= height || row % lines_per_chunk != 0)
            return fail(BMFR_IO_ERR_FORMAT, "%s: chunk %d starts at scanline %d", path, k, y);
        const int y0 = (int)row;
        const int lines = std::min(lines_per_chunk, height - y0);
        const size_t expect = line_bytes * (size_t)lines;
        const unsigned char* src = file.data() + c.at;
        if ((size_t)size == expect) {  // stored raw (always for NONE; for the others when compression did not help)
            raw.assign(src, src + expect);
        } else if (h.compression == COMP_NONE) {
            return fail(BMFR_IO_ERR_FORMAT, "%s: chunk %d has %d bytes, expected %zu", path, k, size, expect);
        } else if (h.compression == COMP_PIZ) {
            if (!piz_decode(src, (size_t)size, h.channels, width, lines, raw) || raw.size() != expect)
                return fail(BMFR_IO_ERR_FORMAT, "%s: corrupt PIZ chunk %d", path, k);
        } else {
            if (h.compression == COMP_RLE) {
                if (!rle_decode(src, (size_t)size, raw, expect)) return fail(BMFR_IO_ERR_FORMAT, "%s: corrupt RLE chunk %d", path, k);
            } else {
                raw.resize(expect);
                uLongf got = (uLongf)expect;
                if (uncompress(raw.data(), &got, src, (uLong)size) != Z_OK || got != expect)
                    return fail(BMFR_IO_ERR_FORMAT, "%s: corrupt ZIP chunk %d", path, k);
            }
            unpredict_and_interleave(raw, tmp);
        }
        for (int l = 0; l < lines; ++l) {
            const unsigned char* line = raw.data() + (size_t)l * line_bytes;
            float* dst = rgb + (size_t)(y0 + l) * width * 3;
            seen[(size_t)(y0 + l)] = 1;
            for (int comp = 0; comp < 3; ++comp) {
                const Channel& ch = h.channels[(size_t)order[comp]];
                const unsigned char* p = line + ch_offset[(size_t)order[comp]];
                if (ch.type == PIXEL_HALF) {
                    for (int x = 0; x < width; ++x) dst[x * 3 + comp] = half_to_float((uint16_t)(p[2 * x] | (p[2 * x + 1] << 8)));
                } else {
                    for (int x = 0; x < width; ++x) {
                        const uint32_t u = (uint32_t)p[4 * x] | ((uint32_t)p[4 * x + 1] << 8) | ((uint32_t)p[4 * x + 2] << 16) |
                                           ((uint32_t)p[4 * x + 3] << 24);
                        memcpy(&dst[x * 3 + comp], &u, 4);
                    }
                }
            }
        }
    }
    for (int y = 0; y < height; ++y)
        if (!seen[(size_t)y]) return fail(BMFR_IO_ERR_FORMAT, "%s: scanline %d is missing", path, y);
    return BMFR_IO_OK;
}

int bmfr_io_parse_camera_header(const char* path, int max_frames, float* matrices, float* offsets, int* n_matrices,
                                int* n_offsets, float* position_limit_squared, float* normal_limit_squared) {
    if (!path || max_frames < 0 || (max_frames > 0 && (!matrices || !offsets)))
        return fail(BMFR_IO_ERR_ARGUMENT, "bmfr_io_parse_camera_header: bad argument");
    std::vector<unsigned char> file;
    if (!read_file(path, file)) return fail(BMFR_IO_ERR_OPEN, "%s: cannot open", path);
    const std::string text = strip_comments(std::string(file.begin(), file.end()));
    std::vector<float> v;
    if (!initialiser_values(text, "camera_matrices", v) || v.empty() || v.size() % 16 != 0)
        return fail(BMFR_IO_ERR_FORMAT, "%s: camera_matrices[][4][4] not found or not a multiple of 16 values (%zu)", path, v.size());
    const int nm = (int)(v.size() / 16);
    for (int i = 0; i < std::min(nm, max_frames) * 16; ++i) matrices[i] = v[(size_t)i];
    if (n_matrices) *n_matrices = nm;
    if (!initialiser_values(text, "pixel_offsets", v) || v.empty() || v.size() % 2 != 0)
        return fail(BMFR_IO_ERR_FORMAT, "%s: pixel_offsets[][2] not found or odd number of values (%zu)", path, v.size());
    const int no = (int)(v.size() / 2);
    for (int i = 0; i < std::min(no, max_frames) * 2; ++i) offsets[i] = v[(size_t)i];
    if (n_offsets) *n_offsets = no;
    if (position_limit_squared && initialiser_values(text, "position_limit_squared", v) && !v.empty()) *position_limit_squared = v[0];
    if (normal_limit_squared && initialiser_values(text, "normal_limit_squared", v) && !v.empty()) *normal_limit_squared = v[0];
    return BMFR_IO_OK;
}

int bmfr_io_psnr(const float* a, const float* b, size_t n, float peak, double* psnr_db) {
    if (!a || !b || !psnr_db || n == 0 || !(peak > 0.f)) return fail(BMFR_IO_ERR_ARGUMENT, "bmfr_io_psnr: bad argument");
    double se = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double d = (double)a[i] - (double)b[i];
        if (d != d) return fail(BMFR_IO_ERR_FORMAT, "bmfr_io_psnr: NaN at element %zu", i);
        se += d * d;
    }
    const double mse = se / (double)n;
    *psnr_db = mse == 0.0 ? HUGE_VAL : 10.0 * log10((double)peak * peak / mse);
    return BMFR_IO_OK;
}

int bmfr_io_ssim_rgb(const float* a, const float* b, int width, int height, float peak, double* ssim) {
    constexpr int R = 5, WIN = 2 * R + 1;
    if (!a || !b || !ssim || width < WIN || height < WIN || !(peak > 0.f))
        return fail(BMFR_IO_ERR_ARGUMENT, "bmfr_io_ssim_rgb: bad argument (images must be at least 11 x 11)");
    double g[WIN], gs = 0.0;
    for (int i = 0; i < WIN; ++i) gs += g[i] = exp(-(double)((i - R) * (i - R)) / (2.0 * 1.5 * 1.5));
    for (int i = 0; i < WIN; ++i) g[i] /= gs;
    const double c1 = (0.01 * peak) * (0.01 * peak), c2 = (0.03 * peak) * (0.03 * peak);
    const int ow = width - 2 * R, oh = height - 2 * R;
    // separable Gaussian moments: five planes (a, b, a^2, b^2, ab), filtered along x then along y
    std::vector<double> tmp((size_t)5 * ow * height), row(5);
    double total = 0.0;
    for (int ch = 0; ch < 3; ++ch) {
        for (int y = 0; y < height; ++y)
            for (int x = 0; x < ow; ++x) {
                double m[5] = {0, 0, 0, 0, 0};
                for (int k = 0; k < WIN; ++k) {
                    const size_t at = ((size_t)y * width + x + k) * 3 + ch;
                    const double va = a[at], vb = b[at];
                    if (va != va || vb != vb) return fail(BMFR_IO_ERR_FORMAT, "bmfr_io_ssim_rgb: NaN at pixel (%d, %d)", x + k, y);
                    m[0] += g[k] * va; m[1] += g[k] * vb; m[2] += g[k] * va * va; m[3] += g[k] * vb * vb; m[4] += g[k] * va * vb;
                }
                for (int q = 0; q < 5; ++q) tmp[((size_t)q * height + y) * ow + x] = m[q];
            }
        double sum = 0.0;
        for (int y = 0; y < oh; ++y)
            for (int x = 0; x < ow; ++x) {
                double m[5] = {0, 0, 0, 0, 0};
                for (int k = 0; k < WIN; ++k)
                    for (int q = 0; q < 5; ++q) m[q] += g[k] * tmp[((size_t)q * height + y + k) * ow + x];
                const double va = m[2] - m[0] * m[0], vb = m[3] - m[1] * m[1], cov = m[4] - m[0] * m[1];
                sum += ((2.0 * m[0] * m[1] + c1) * (2.0 * cov + c2)) / ((m[0] * m[0] + m[1] * m[1] + c1) * (va + vb + c2));
            }
        total += sum / ((double)ow * oh);
    }
    *ssim = total / 3.0;
    return BMFR_IO_OK;
}

void bmfr_io_tone_map(float* rgb, size_t n) {
    if (!rgb) return;
    for (size_t i = 0; i < n; ++i) {
        float v = rgb[i] > 0.f ? rgb[i] : 0.f;  // max(0, v); NaN -> 0
        v = powf(v, 0.454545f);
        rgb[i] = v < 1.f ? v : 1.f;
    }
}

int bmfr_io_write_png_rgb(const char* path, int width, int height, const float* rgb, size_t row_stride_floats) {
    if (!path || !rgb || width <= 0 || height <= 0 || row_stride_floats < (size_t)width * 3)
        return fail(BMFR_IO_ERR_ARGUMENT, "bmfr_io_write_png_rgb: bad argument");
    std::vector<unsigned char> raw((size_t)height * ((size_t)width * 3 + 1));
    for (int y = 0; y < height; ++y) {
        unsigned char* row = raw.data() + (size_t)y * ((size_t)width * 3 + 1);
        *row++ = 0;  // filter type: none
        const float* src = rgb + (size_t)y * row_stride_floats;
        for (int i = 0; i < width * 3; ++i) {
            float v = src[i];
            v = v > 0.f ? (v < 1.f ? v : 1.f) : 0.f;  // NaN fails v > 0 -> 0
            row[i] = (unsigned char)(v * 255.f + 0.5f);
        }
    }
    uLongf zn = compressBound((uLong)raw.size());
    std::vector<unsigned char> z(zn);
    if (compress2(z.data(), &zn, raw.data(), (uLong)raw.size(), 6) != Z_OK) return fail(BMFR_IO_ERR_FORMAT, "%s: zlib failed", path);
    std::vector<unsigned char> png = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<unsigned char> ihdr;
    put_be32(ihdr, (uint32_t)width);
    put_be32(ihdr, (uint32_t)height);
    const unsigned char tail[5] = {8, 2, 0, 0, 0};  // 8 bits, colour type 2 (RGB), deflate, adaptive filtering, no interlace
    ihdr.insert(ihdr.end(), tail, tail + 5);
    put_chunk(png, "IHDR", ihdr.data(), ihdr.size());
    put_chunk(png, "IDAT", z.data(), (size_t)zn);
    put_chunk(png, "IEND", nullptr, 0);
    FILE* f = fopen(path, "wb");
    if (!f) return fail(BMFR_IO_ERR_OPEN, "%s: cannot create", path);  // bmfr.cpp:541-546
    const bool ok = fwrite(png.data(), 1, png.size(), f) == png.size();
    if (fclose(f) != 0 || !ok) return fail(BMFR_IO_ERR_OPEN, "%s: write failed", path);
    return BMFR_IO_OK;
}

}  // extern "C"
