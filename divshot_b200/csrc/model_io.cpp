// model_io.cpp — Gaussian-model writers and readers behind include/dvs_model_io.h.
//
// Host-only C++ (compiled by g++ with -ffp-contract=off: the quantisers below are literal float/double operation
// sequences so that every written byte equals what the reference's writers produce on the same input).
// Formats and the reference code that defines them:
//   PLY            external/tinygsplat/tiny_gsplat.cpp:168-241   (reader :632-722)
//   .splat         tiny_gsplat.cpp:243-291                       (reader :724-768)
//   compressed PLY tiny_gsplat.cpp:293-395, tiny_gsplat.hpp:331-468   (reader hpp:470-534)
//   .dvsplat       tiny_gsplat.cpp:994-1117                      (reader :1119-1191)
//   .spz           tiny_gsplat.cpp:1243-1272 -> external/spz/src/load-spz.cc:216-333,533-546,598-607 (v3, gzip)
// Design: every writer is "quantise SoA inputs straight into ONE output buffer, one write() at the end"; the two
// chunked formats share a Morton ordering + a chunk-bounds helper.  The readers return rows in the reference's
// RichPoint layout (59 floats) so they can be compared value for value with the reference readers.
//
// Reference quirks kept on purpose (byte parity; DESIGN.md §7 lists them):
//   Q1 chunk bounds are seeded with the element at raw position `start`, not order[start] (tiny_gsplat.hpp:332);
//   Q2 .dvsplat attribute blocks are addressed through the Morton permutation of the *degree list position*
//      (tiny_gsplat.cpp:1082) — identical to the position order only when all degrees are equal;
//   Q3 .spz SH coefficients are stored at [p*15+j+c] (tiny_gsplat.cpp:1262-1267) unless DVS_IO_SPZ_SH_FIXED;
//   Q4 the PLY reader of a 2DGS file (no scale_2) leaves scale.xy = 0 and sets scale.z = log(1e-6).
#include "dvs_model_io.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return -1; }

constexpr float kShC0 = 0.28209479177387814f;
constexpr int kRow = DVS_IO_ROW_FLOATS;  // 3 + 48 + 1 + 3 + 4
constexpr int kRowOpacity = 51, kRowScale = 52, kRowRot = 55;

struct Cloud {  // the trainer's tensors, borrowed
    int64_t N;
    const float *pos, *sh0, *shN, *opac, *scale, *rot;
    const uint8_t* deg;
};

// ---------------------------------------------------------------------------------------------- quantisers
inline float clampf(float x, float lo, float hi) {  // min(max(x, lo), hi) with `<` only (NaN passes through)
    const float a = (x < lo) ? lo : x;
    return (hi < a) ? hi : a;
}
inline float sigmoidf(float x) { return 1 / (1 + std::exp(-x)); }
inline uint8_t round_u8(float x) { return static_cast<uint8_t>(std::clamp(std::round(x), 0.0f, 255.0f)); }

// n-bit unorm: floor(value*(2^bits-1) + 0.5) where the product is float and the sum/floor are double
inline uint32_t unorm(float value, int bits) {
    const int t = (1 << bits) - 1;
    const double r = std::floor(value * t + 0.5);
    return static_cast<uint32_t>(clampf(static_cast<float>(r), 0.f, static_cast<float>(t)));
}
inline uint32_t pack_11_10_11(float x, float y, float z) { return unorm(x, 11) << 21 | unorm(y, 10) << 11 | unorm(z, 11); }
inline float span01(float x, float lo, float hi) { return (hi - lo < 0.00001) ? 0 : (x - lo) / (hi - lo); }

struct Quat { float c[4]; };
// v * (1/sqrt((v0²+v1²)+(v2²+v3²)))   — the pairwise dot product and reciprocal-sqrt scaling the PLY-family uses
inline Quat unit_pairwise(const float* v) {
    const float d = (v[0] * v[0] + v[1] * v[1]) + (v[2] * v[2] + v[3] * v[3]);
    const float s = 1.0f / std::sqrt(d);
    return {{v[0] * s, v[1] * s, v[2] * s, v[3] * s}};
}
// v / sqrt(v0²+v1²+v2²+v3²)           — the left-to-right sum and per-component division .spz uses
inline Quat unit_sequential(const float* v) {
    const float n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
    return {{v[0] / n, v[1] / n, v[2] / n, v[3] / n}};
}

// 2+10+10+10 "largest component dropped" quaternion of the compressed PLY
inline uint32_t pack_quat_2_10_10_10(const float* wxyz) {
    Quat q = unit_pairwise(wxyz);
    int big = 0;
    for (int i = 1; i < 4; i++)
        if (std::abs(q.c[big]) < std::abs(q.c[i])) big = i;
    if (q.c[big] < 0)
        for (float& v : q.c) v = -v;
    const float norm = static_cast<float>(std::sqrt(2.0) * 0.5f);
    uint32_t out = static_cast<uint32_t>(big);
    for (int i = 0; i < 4; i++)
        if (i != big) out = (out << 10) | unorm(q.c[i] * norm + 0.5f, 10);
    return out;
}
inline uint32_t pack_rgba_8888(const float* dc, float logit) {
    return unorm(dc[0] * kShC0 + 0.5f, 8) << 24 | unorm(dc[1] * kShC0 + 0.5f, 8) << 16 | unorm(dc[2] * kShC0 + 0.5f, 8) << 8 |
           unorm(sigmoidf(logit), 8);
}
// 8-bit SH bucket quantiser shared by .dvsplat and .spz (0 always maps to a bucket centre)
inline uint8_t quant_sh(float x, int bucket) {
    int q = static_cast<int>(std::round(x * 128.0f) + 128.0f);
    q = (q + bucket / 2) / bucket * bucket;
    return static_cast<uint8_t>(std::clamp(q, 0, 255));
}
inline float dequant_sh(uint8_t v) { return (static_cast<float>(v) - 128.0f) / 128.0f; }
inline float logit_of(float p) { return std::log(p / (1.0f - p)); }
constexpr float kWideColour = 0.15f;  // .dvsplat/.spz DC colour scale

// float -> int32 the way `cvttss2si` does it (the reference leans on that for degenerate extents: NaN -> INT_MIN)
inline int32_t trunc_i32(float f) {
    if (!(f > -2147483904.0f && f < 2147483648.0f)) return std::numeric_limits<int32_t>::min();
    return static_cast<int32_t>(f);
}

// ---------------------------------------------------------------------------------------------- chunk helpers
struct Box { float lo[3], hi[3]; };

// Morton (z-order) permutation over the global bounding box, 21 bits per axis, x in the lowest bit of each triple.
std::vector<uint64_t> morton_order(const float* pos, int64_t N) {
    Box g;
    for (int a = 0; a < 3; a++) g.lo[a] = g.hi[a] = pos[a];
    for (int64_t i = 0; i < N; i++)
        for (int a = 0; a < 3; a++) {
            const float v = pos[3 * i + a];
            if (v < g.lo[a]) g.lo[a] = v;
            if (g.hi[a] < v) g.hi[a] = v;
        }
    std::vector<std::pair<uint64_t, int>> keyed(static_cast<size_t>(N));
    for (int64_t i = 0; i < N; i++) {
        uint64_t code = 0;
        for (int a = 0; a < 3; a++) {
            const float rel = (pos[3 * i + a] - g.lo[a]) / (g.hi[a] - g.lo[a]);
            const uint32_t cell = static_cast<uint32_t>(trunc_i32(2097151.0f * rel)) & 0x1FFFFFu;
            for (int b = 0; b < 21; b++) code |= static_cast<uint64_t>((cell >> b) & 1u) << (3 * b + a);
        }
        keyed[static_cast<size_t>(i)] = {code, static_cast<int>(i)};
    }
    // same algorithm, comparator and element type as the reference's call (tiny_gsplat.cpp:322-325), so that even
    // equal codes land in the same order
    std::sort(keyed.begin(), keyed.end(),
              [](const std::pair<uint64_t, int>& a, const std::pair<uint64_t, int>& b) { return a.first < b.first; });
    std::vector<uint64_t> order(static_cast<size_t>(N));
    for (int64_t i = 0; i < N; i++) order[static_cast<size_t>(i)] = static_cast<uint64_t>(keyed[static_cast<size_t>(i)].second);
    return order;
}

// bounds of v[order[j]] for j in [start, end) — seeded with v[start] (quirk Q1)
Box chunk_box(const float* v, const std::vector<uint64_t>& order, size_t start, size_t end) {
    Box b;
    for (int a = 0; a < 3; a++) b.lo[a] = b.hi[a] = v[3 * start + a];
    for (size_t j = start; j < std::min(end, order.size()); j++)
        for (int a = 0; a < 3; a++) {
            const float x = v[3 * order[j] + a];
            if (x < b.lo[a]) b.lo[a] = x;
            if (b.hi[a] < x) b.hi[a] = x;
        }
    return b;
}
inline uint32_t pack_in_box(const float* v, const Box& b) {
    return pack_11_10_11(span01(v[0], b.lo[0], b.hi[0]), span01(v[1], b.lo[1], b.hi[1]), span01(v[2], b.lo[2], b.hi[2]));
}

struct Bytes {
    std::vector<uint8_t> buf;
    explicit Bytes(size_t n = 0) : buf(n) {}
    template <class T> void put(size_t off, T v) { std::memcpy(buf.data() + off, &v, sizeof(T)); }
    template <class T> T get(size_t off) const { T v; std::memcpy(&v, buf.data() + off, sizeof(T)); return v; }
};

bool flush(const std::string& path, const std::string& header, const uint8_t* body, size_t n) {
    std::ofstream out(path, std::ios::binary);
    if (!out.good()) return false;
    out.write(header.data(), static_cast<std::streamsize>(header.size()));
    if (n) out.write(reinterpret_cast<const char*>(body), static_cast<std::streamsize>(n));
    out.close();
    return out.good();
}

// ---------------------------------------------------------------------------------------------- writers
int write_ply(const std::string& path, const Cloud& c, uint32_t flags) {
    std::string h = "ply\nformat binary_little_endian 1.0\ncomment generated by spaltX\n";  // (sic) tiny_gsplat.cpp:188
    if (flags & DVS_IO_ANTIALIASED) h += "comment splatx.anti_aliasing=1\n";
    h += "element vertex " + std::to_string(c.N) + "\nproperty float x\nproperty float y\nproperty float z\n";
    for (int i = 0; i < 3; i++) h += "property float f_dc_" + std::to_string(i) + "\n";
    for (int i = 0; i < 45; i++) h += "property float f_rest_" + std::to_string(i) + "\n";
    h += "property float opacity\n";
    for (int i = 0; i < 3; i++) h += "property float scale_" + std::to_string(i) + "\n";
    for (int i = 0; i < 4; i++) h += "property float rot_" + std::to_string(i) + "\n";
    h += "end_header\n";
    std::vector<float> rows(static_cast<size_t>(c.N) * kRow);
    for (int64_t i = 0; i < c.N; i++) {
        float* r = rows.data() + static_cast<size_t>(i) * kRow;
        std::memcpy(r, c.pos + 3 * i, 12);
        std::memcpy(r + 3, c.sh0 + 3 * i, 12);
        const float* rest = c.shN + 45 * i;  // [15][3] -> three planes of 15 (tiny_gsplat.cpp:231-236)
        for (int j = 0; j < 15; j++)
            for (int ch = 0; ch < 3; ch++) r[6 + ch * 15 + j] = rest[3 * j + ch];
        r[kRowOpacity] = c.opac[i];
        std::memcpy(r + kRowScale, c.scale + 3 * i, 12);
        std::memcpy(r + kRowRot, c.rot + 4 * i, 16);
    }
    return flush(path, h, reinterpret_cast<const uint8_t*>(rows.data()), rows.size() * 4) ? 0 : fail("cannot write " + path);
}

int write_splat(const std::string& path, const Cloud& c) {
    Bytes out(static_cast<size_t>(c.N) * 32);
    for (int64_t i = 0; i < c.N; i++) {
        const size_t o = static_cast<size_t>(i) * 32;
        for (int a = 0; a < 3; a++) out.put<float>(o + 4 * a, c.pos[3 * i + a]);
        for (int a = 0; a < 3; a++) out.put<float>(o + 12 + 4 * a, std::exp(c.scale[3 * i + a]));
        for (int a = 0; a < 3; a++) {  // colour in double precision, clamped as float, truncated
            const double v = (0.5 + 0.28209479177387814 * c.sh0[3 * i + a]) * 255;
            out.buf[o + 24 + a] = static_cast<uint8_t>(clampf(static_cast<float>(v), 0.f, 255.f));
        }
        out.buf[o + 27] = static_cast<uint8_t>(clampf((1 / (1 + std::exp(-c.opac[i]))) * 255, 0.f, 255.f));
        const Quat q = unit_pairwise(c.rot + 4 * i);
        for (int a = 0; a < 4; a++) out.buf[o + 28 + a] = static_cast<uint8_t>(std::clamp<float>(q.c[a] * 128 + 128, 0, 255));
    }
    return flush(path, "", out.buf.data(), out.buf.size()) ? 0 : fail("cannot write " + path);
}

int write_compressed_ply(const std::string& path, const Cloud& c, uint32_t flags) {
    const size_t N = static_cast<size_t>(c.N), chunks = (N + 255) / 256;
    const std::vector<uint64_t> order = morton_order(c.pos, c.N);
    Bytes out(chunks * 48 + N * 16);
    const size_t vertex0 = chunks * 48;
    for (size_t k = 0; k < chunks; k++) {
        const size_t start = k * 256, end = start + 256;
        const Box pb = chunk_box(c.pos, order, start, end), sb = chunk_box(c.scale, order, start, end);
        for (int a = 0; a < 3; a++) {
            out.put<float>(k * 48 + 4 * a, pb.lo[a]);
            out.put<float>(k * 48 + 12 + 4 * a, pb.hi[a]);
            out.put<float>(k * 48 + 24 + 4 * a, sb.lo[a]);
            out.put<float>(k * 48 + 36 + 4 * a, sb.hi[a]);
        }
        for (size_t j = start; j < std::min(end, N); j++) {
            const size_t i = order[j], o = vertex0 + j * 16;
            out.put<uint32_t>(o, pack_in_box(c.pos + 3 * i, pb));
            out.put<uint32_t>(o + 4, pack_quat_2_10_10_10(c.rot + 4 * i));
            out.put<uint32_t>(o + 8, pack_in_box(c.scale + 3 * i, sb));
            out.put<uint32_t>(o + 12, pack_rgba_8888(c.sh0 + 3 * i, c.opac[i]));
        }
    }
    std::string h = "ply\nformat binary_little_endian 1.0\ncomment generated by diverseshot\n";
    if (flags & DVS_IO_ANTIALIASED) h += "comment splatx.anti_aliasing=1\n";
    h += "element chunk " + std::to_string(chunks) + "\n";
    for (const char* lim : {"min", "max"})
        for (const char* ax : {"x", "y", "z"}) h += std::string("property float ") + lim + "_" + ax + "\n";
    for (const char* lim : {"min", "max"})
        for (const char* ax : {"x", "y", "z"}) h += std::string("property float ") + lim + "_scale_" + ax + "\n";
    h += "element vertex " + std::to_string(N) + "\n";
    for (const char* p : {"position", "rotation", "scale", "color"}) h += std::string("property uint packed_") + p + "\n";
    h += "end_header\n";
    return flush(path, h, out.buf.data(), out.buf.size()) ? 0 : fail("cannot write " + path);
}

int write_dvsplat(const std::string& path, const Cloud& c) {
    const size_t N = static_cast<size_t>(c.N), chunks = (N + 255) / 256;
    const std::vector<uint64_t> order = morton_order(c.pos, c.N);
    std::vector<int> by_degree[4];
    size_t attr_bytes = 0;
    for (int d = 0; d < 4; d++) {
        for (size_t i = 0; i < N; i++)
            if ((c.deg ? c.deg[i] : 3) == d) by_degree[d].push_back(static_cast<int>(i));
        attr_bytes += by_degree[d].size() * static_cast<size_t>(7 + 3 * (d + 1) * (d + 1));
    }
    const size_t header = 28, bounds0 = header, pos0 = header + chunks * 24, attr0 = pos0 + N * 4;
    Bytes out(attr0 + attr_bytes);
    out.put<uint32_t>(0, static_cast<uint32_t>(N));
    out.put<uint32_t>(4, static_cast<uint32_t>(chunks));
    for (int d = 0; d < 4; d++) out.put<uint32_t>(8 + 4 * d, static_cast<uint32_t>(by_degree[d].size()));
    out.put<uint32_t>(24, 0u);
    for (size_t k = 0; k < chunks; k++) {
        const size_t start = k * 256, end = start + 256;
        const Box pb = chunk_box(c.pos, order, start, end);
        for (int a = 0; a < 3; a++) {
            out.put<float>(bounds0 + k * 24 + 4 * a, pb.lo[a]);
            out.put<float>(bounds0 + k * 24 + 12 + 4 * a, pb.hi[a]);
        }
        for (size_t j = start; j < std::min(end, N); j++) out.put<uint32_t>(pos0 + j * 4, pack_in_box(c.pos + 3 * order[j], pb));
    }
    size_t off = attr0;
    for (int d = 0; d < 4; d++) {
        const int ncoef = (d + 1) * (d + 1) - 1;
        const size_t stride = static_cast<size_t>(7 + 3 * (1 + ncoef));
        for (size_t s = 0; s < by_degree[d].size(); s++) {
            const size_t i = order[static_cast<size_t>(by_degree[d][s])];  // quirk Q2
            uint8_t* r = out.buf.data() + off + s * stride;
            for (int a = 0; a < 3; a++) r[a] = round_u8((c.scale[3 * i + a] + 10.0f) * 16.0f);
            Quat q = unit_pairwise(c.rot + 4 * i);
            const float sgn = q.c[0] < 0 ? -127.5f : 127.5f;
            for (int a = 1; a < 4; a++) r[2 + a] = round_u8(q.c[a] * sgn + 127.5f);
            r[6] = round_u8(sigmoidf(c.opac[i]) * 255.0f);
            for (int a = 0; a < 3; a++) r[7 + a] = round_u8(c.sh0[3 * i + a] * (kWideColour * 255.0f) + (0.5f * 255.0f));
            const float* rest = c.shN + 45 * i;
            for (int j = 0; j < ncoef * 3; j++) r[10 + j] = quant_sh(rest[j], j < 9 ? 8 : 16);
        }
        off += by_degree[d].size() * stride;
    }
    return flush(path, "", out.buf.data(), out.buf.size()) ? 0 : fail("cannot write " + path);
}

bool gzip(const std::vector<uint8_t>& raw, std::vector<uint8_t>* out) {
    z_stream zs = {};
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 16 + MAX_WBITS, 9, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    out->resize(deflateBound(&zs, static_cast<uLong>(raw.size())));
    zs.next_in = const_cast<Bytef*>(raw.data());
    zs.avail_in = static_cast<uInt>(raw.size());
    zs.next_out = out->data();
    zs.avail_out = static_cast<uInt>(out->size());
    const int rc = deflate(&zs, Z_FINISH);
    out->resize(zs.total_out);
    deflateEnd(&zs);
    return rc == Z_STREAM_END;
}
bool gunzip(const std::vector<uint8_t>& in, std::vector<uint8_t>* out) {
    z_stream zs = {};
    if (inflateInit2(&zs, 16 | MAX_WBITS) != Z_OK) return false;
    zs.next_in = const_cast<Bytef*>(in.data());
    zs.avail_in = static_cast<uInt>(in.size());
    out->clear();
    std::vector<uint8_t> block(1 << 16);
    int rc = Z_OK;
    while (rc == Z_OK) {
        zs.next_out = block.data();
        zs.avail_out = static_cast<uInt>(block.size());
        rc = inflate(&zs, Z_NO_FLUSH);
        if (rc != Z_OK && rc != Z_STREAM_END) break;
        out->insert(out->end(), block.data(), block.data() + (block.size() - zs.avail_out));
    }
    inflateEnd(&zs);
    return rc == Z_STREAM_END;
}

constexpr uint32_t kSpzMagic = 0x5053474e;  // "NGSP"
constexpr float kInvSqrt2 = static_cast<float>(0.707106781186547524401);

int write_spz(const std::string& path, const Cloud& c, uint32_t flags) {
    const size_t N = static_cast<size_t>(c.N);
    if (N >= (size_t{1} << 31) / 64) return fail(".spz: too many points for a single gzip stream");
    // section offsets: header 16 | positions 9N | alphas N | colours 3N | scales 3N | rotations 4N | sh 45N
    const size_t p0 = 16, a0 = p0 + 9 * N, c0 = a0 + N, s0 = c0 + 3 * N, r0 = s0 + 3 * N, h0 = r0 + 4 * N;
    Bytes raw(h0 + 45 * N);
    raw.put<uint32_t>(0, kSpzMagic);
    raw.put<uint32_t>(4, 3u);
    raw.put<uint32_t>(8, static_cast<uint32_t>(N));
    raw.buf[12] = 3;   // SH degree
    raw.buf[13] = 12;  // fractional bits of the 24-bit fixed-point coordinates
    raw.buf[14] = (flags & DVS_IO_ANTIALIASED) ? 1 : 0;
    raw.buf[15] = 0;
    // the SH block as the reference's caller lays it out before packing (quirk Q3)
    std::vector<float> sh(45 * N, 0.f);
    for (size_t p = 0; p < N; p++)
        for (int j = 0; j < 15; j++)
            for (int ch = 0; ch < 3; ch++) {
                const size_t dst = (flags & DVS_IO_SPZ_SH_FIXED) ? (p * 15 + j) * 3 + ch : (p * 15 + j) + ch;
                sh[dst] = c.shN[45 * p + 3 * j + ch];
            }
    for (size_t i = 0; i < 3 * N; i++) {
        const int32_t fx = static_cast<int32_t>(std::round(1.0f * c.pos[i] * 4096.0f));
        raw.buf[p0 + 3 * i] = fx & 0xff;
        raw.buf[p0 + 3 * i + 1] = (fx >> 8) & 0xff;
        raw.buf[p0 + 3 * i + 2] = (fx >> 16) & 0xff;
        raw.buf[s0 + i] = round_u8((c.scale[i] + 10.0f) * 16.0f);
        raw.buf[c0 + i] = round_u8(c.sh0[i] * (kWideColour * 255.0f) + (0.5f * 255.0f));
    }
    for (size_t i = 0; i < N; i++) {
        raw.buf[a0 + i] = round_u8(sigmoidf(c.opac[i]) * 255.0f);
        const float xyzw[4] = {c.rot[4 * i + 1], c.rot[4 * i + 2], c.rot[4 * i + 3], c.rot[4 * i]};
        const Quat q = unit_sequential(xyzw);
        unsigned big = 0;
        for (unsigned k = 1; k < 4; k++)
            if (std::abs(q.c[k]) > std::abs(q.c[big])) big = k;
        const unsigned negate = q.c[big] < 0;
        uint32_t comp = big;
        for (unsigned k = 0; k < 4; k++)
            if (k != big) {
                const uint32_t neg = static_cast<uint32_t>(q.c[k] < 0) ^ negate;
                const uint32_t mag = static_cast<uint32_t>(511.0f * (std::fabs(q.c[k]) / kInvSqrt2) + 0.5f);
                comp = (comp << 10) | (neg << 9) | mag;
            }
        raw.put<uint32_t>(r0 + 4 * i, comp);
    }
    for (size_t p = 0; p < N; p++)
        for (int j = 0; j < 45; j++) raw.buf[h0 + 45 * p + j] = quant_sh(1.0f * sh[45 * p + j], j < 9 ? 8 : 16);
    std::vector<uint8_t> gz;
    if (!gzip(raw.buf, &gz)) return fail(".spz: deflate failed");
    return flush(path, "", gz.data(), gz.size()) ? 0 : fail("cannot write " + path);
}

// ---------------------------------------------------------------------------------------------- readers
// Reduced PLY as GaussianModel::save_to_file reaches it (gaussian_model.cpp:445-448: the codebook / quantised / half-float
// arguments keep their defaults, tiny_gsplat.hpp:625-636), i.e. the float branch of tiny_gsplat.cpp:398-630: the Gaussians
// are grouped by SH degree into four `element vertex` blocks (input order inside a block) and a row of degree d is
//   pos[3] | f_dc[3] | f_rest[3 ((d+1)^2 - 1)] | opacity | scale[3] | rot[4]        (all float32)
// Quirk Q8, reproduced for byte parity: coefficient j is copied from &shs_n[i][j] — the 12 bytes starting at FLOAT j of the
// 45-float row, i.e. the overlapping window [j, j+2] instead of [3j, 3j+2] (tiny_gsplat.cpp:533-534); the reader returns
// exactly those windows.  DVS_IO_REDUCED_SH_FIXED writes the intended coefficients.
int write_reduced_ply(const std::string& path, const Cloud& c, uint32_t flags) {
    std::vector<int64_t> ids[4];
    for (int64_t i = 0; i < c.N; i++) {
        const int d = c.deg ? c.deg[i] : 3;
        if (d >= 0 && d <= 3) ids[d].push_back(i);  // other degree values are in no block (as in the reference)
    }
    std::string h = "ply\nformat binary_little_endian 1.0\ncomment generated by diverseshot\n";
    size_t bytes = 0;
    for (int d = 0; d < 4; d++) {
        const int coeffs = (d + 1) * (d + 1) - 1;
        h += "element vertex " + std::to_string(ids[d].size()) + "\nproperty float x\nproperty float y\nproperty float z\n";
        h += "property float f_dc_0\nproperty float f_dc_1\nproperty float f_dc_2\n";
        for (int j = 0; j < coeffs; j++)
            for (int ch = 0; ch < 3; ch++) h += "property float f_rest_" + std::to_string(3 * j + ch) + "\n";
        h += "property float opacity\nproperty float scale_0\nproperty float scale_1\nproperty float scale_2\n";
        h += "property float rot_0\nproperty float rot_1\nproperty float rot_2\nproperty float rot_3\n";
        bytes += ids[d].size() * static_cast<size_t>(4 * (3 + 3 + 3 * coeffs + 1 + 3 + 4));
    }
    h += "end_header\n";
    std::vector<float> body(bytes / 4);
    float* w = body.data();
    for (int d = 0; d < 4; d++) {
        const int coeffs = (d + 1) * (d + 1) - 1;
        for (const int64_t i : ids[d]) {
            std::memcpy(w, c.pos + 3 * i, 12); w += 3;
            std::memcpy(w, c.sh0 + 3 * i, 12); w += 3;
            const float* rest = c.shN + 45 * i;
            for (int j = 0; j < coeffs; j++, w += 3)
                std::memcpy(w, (flags & DVS_IO_REDUCED_SH_FIXED) ? rest + 3 * j : rest + j, 12);
            *w++ = c.opac[i];
            std::memcpy(w, c.scale + 3 * i, 12); w += 3;
            std::memcpy(w, c.rot + 4 * i, 16); w += 4;
        }
    }
    return flush(path, h, reinterpret_cast<const uint8_t*>(body.data()), bytes) ? 0 : fail("cannot write " + path);
}

bool slurp(const std::string& path, std::vector<uint8_t>* out) {
    std::ifstream in(path, std::ios::binary | std::ios::ate);
    if (!in.good()) return false;
    out->resize(static_cast<size_t>(in.tellg()));
    in.seekg(0, std::ios::beg);
    if (!out->empty()) in.read(reinterpret_cast<char*>(out->data()), static_cast<std::streamsize>(out->size()));
    return in.good() || in.eof();
}

struct PlyHeader {
    uint64_t vertices = 0, chunks = 0;
    uint32_t stride = 0;
    bool antialiased = false, has_scale2 = false;
    std::unordered_map<std::string, uint32_t> offset;
    size_t body = 0;  // byte offset of the binary payload
};
// The reference's "parser": line-based, property offsets accumulate only once `element vertex` has been seen.
bool parse_ply_header(const std::vector<uint8_t>& f, PlyHeader* h) {
    size_t p = 0;
    bool in_vertex = false;
    while (p < f.size()) {
        const uint8_t* nl = static_cast<const uint8_t*>(std::memchr(f.data() + p, '\n', f.size() - p));
        const size_t e = nl ? static_cast<size_t>(nl - f.data()) : f.size();
        const std::string line(reinterpret_cast<const char*>(f.data() + p), e - p);
        p = e + 1;
        if (line == "end_header") {
            // the payload starts after the newline that ends the header: a file that stops at "end_header" without one
            // has no payload position (p would be size + 1 and every later size check would wrap)
            if (!nl) { g_err = "PLY: truncated header"; return false; }
            h->body = p;
            return true;
        }
        std::string a, b, name;
        std::istringstream ss(line);
        if (line.find("anti_aliasing=1") != std::string::npos) h->antialiased = true;
        if (line.find("element vertex") != std::string::npos) { ss >> a >> b >> h->vertices; in_vertex = true; }
        else if (line.find("element chunk") != std::string::npos) { ss >> a >> b >> h->chunks; }
        else if (line.find("property") != std::string::npos) {
            ss >> a >> b >> name;
            if (b != "float" && b != "uint") { g_err = "PLY: unsupported property type " + b; return false; }
            h->offset[name] = h->stride;
            if (in_vertex) h->stride += 4;
        }
        if (line.find("scale_2") != std::string::npos) h->has_scale2 = true;
    }
    g_err = "PLY: no end_header";
    return false;
}

int64_t read_ply(const std::vector<uint8_t>& f, float* rows, int64_t cap, uint32_t* flags) {
    PlyHeader h;
    if (!parse_ply_header(f, &h)) return -1;
    if (h.antialiased) *flags |= DVS_IO_ANTIALIASED;
    const int64_t n = static_cast<int64_t>(h.vertices);
    if (n <= 0) return fail("PLY: no vertices");
    if (h.stride == 0 || static_cast<uint64_t>(n) > (f.size() - h.body) / h.stride) return fail("PLY: truncated payload");  // (no product: a forged count cannot wrap it)
    if (!rows) return n;
    auto field = [&](const char* name) -> int64_t {
        auto it = h.offset.find(name);
        return it == h.offset.end() ? -1 : static_cast<int64_t>(it->second);
    };
    int64_t col[kRow];  // source byte offset of each row float, -1 = absent (left 0)
    const char* xyz[3] = {"x", "y", "z"};
    for (int a = 0; a < 3; a++) col[a] = field(xyz[a]);
    for (int a = 0; a < 3; a++) col[3 + a] = field(("f_dc_" + std::to_string(a)).c_str());
    for (int a = 0; a < 45; a++) col[6 + a] = field(("f_rest_" + std::to_string(a)).c_str());
    col[kRowOpacity] = field("opacity");
    for (int a = 0; a < 3; a++) col[kRowScale + a] = h.has_scale2 ? field(("scale_" + std::to_string(a)).c_str()) : -1;
    for (int a = 0; a < 4; a++) col[kRowRot + a] = field(("rot_" + std::to_string(a)).c_str());
    for (int64_t i = 0; i < std::min(n, cap); i++) {
        const uint8_t* src = f.data() + h.body + static_cast<size_t>(i) * h.stride;
        float* r = rows + i * kRow;
        for (int k = 0; k < kRow; k++) {
            r[k] = 0.f;
            if (col[k] >= 0) std::memcpy(r + k, src + col[k], 4);
        }
        if (!h.has_scale2) r[kRowScale + 2] = std::log(1e-6f);  // quirk Q4
    }
    return n;
}

int64_t read_splat(const std::vector<uint8_t>& f, float* rows, int64_t cap) {
    const int64_t n = static_cast<int64_t>(f.size() / 32);
    if (n <= 0) return fail(".splat: empty file");
    if (!rows) return n;
    for (int64_t i = 0; i < std::min(n, cap); i++) {
        const uint8_t* s = f.data() + i * 32;
        float* r = rows + i * kRow;
        std::fill(r, r + kRow, 0.f);
        float v[6];
        std::memcpy(v, s, 24);
        for (int a = 0; a < 3; a++) { r[a] = v[a]; r[kRowScale + a] = std::log(v[3 + a]); }
        for (int a = 0; a < 3; a++) r[3 + a] = static_cast<float>((s[24 + a] / 255.0 - 0.5) / kShC0);
        const float o = s[27] / 255.0f;
        r[kRowOpacity] = std::log(o / (1 - o));
        for (int a = 0; a < 4; a++) r[kRowRot + a] = (s[28 + a] - 128) / 128.0f;
    }
    return n;
}

inline float from_unorm(uint32_t v, int bits) {
    const uint32_t m = (1u << bits) - 1;
    return static_cast<float>(v & m) / m;
}
inline void unpack_in_box(uint32_t v, const float* lo, const float* hi, float* out) {
    const float u[3] = {from_unorm(v >> 21, 11), from_unorm(v >> 11, 10), from_unorm(v, 11)};
    for (int a = 0; a < 3; a++) out[a] = u[a] * (hi[a] - lo[a]) + lo[a];
}

int64_t read_compressed_ply(const std::vector<uint8_t>& f, float* rows, int64_t cap, uint32_t* flags) {
    PlyHeader h;
    if (!parse_ply_header(f, &h)) return -1;
    if (h.antialiased) *flags |= DVS_IO_ANTIALIASED;
    const size_t N = h.vertices, chunks = h.chunks;
    if (N == 0 || chunks == 0) return fail("compressed PLY: no chunks / vertices");
    if (N > (f.size() - h.body) / 16 || chunks > (f.size() - h.body) / 48 || f.size() - h.body < chunks * 48 + N * 16)
        return fail("compressed PLY: truncated payload");
    if (!rows) return static_cast<int64_t>(N);
    const uint8_t* body = f.data() + h.body;
    const double inv_norm = 1.0 / (std::sqrt(2.0) * 0.5);
    for (size_t i = 0; i < std::min<size_t>(N, static_cast<size_t>(cap)); i++) {
        float b[12];
        std::memcpy(b, body + (i / 256) * 48, 48);
        uint32_t w[4];
        std::memcpy(w, body + chunks * 48 + i * 16, 16);
        float* r = rows + i * kRow;
        std::fill(r, r + kRow, 0.f);
        unpack_in_box(w[0], b, b + 3, r);
        unpack_in_box(w[2], b + 6, b + 9, r + kRowScale);
        const float qa = static_cast<float>((from_unorm(w[1] >> 20, 10) - 0.5) * inv_norm);
        const float qb = static_cast<float>((from_unorm(w[1] >> 10, 10) - 0.5) * inv_norm);
        const float qc = static_cast<float>((from_unorm(w[1], 10) - 0.5) * inv_norm);
        const float m = std::sqrt(1.0f - (qa * qa + qb * qb + qc * qc));
        const float rest[3] = {qa, qb, qc};
        const int big = static_cast<int>(w[1] >> 30);
        for (int k = 0, s = 0; k < 4; k++) r[kRowRot + k] = (k == big) ? m : rest[s++];
        for (int a = 0; a < 3; a++) r[3 + a] = (from_unorm(w[3] >> (24 - 8 * a), 8) - 0.5f) / kShC0;
        r[kRowOpacity] = -std::log(1 / from_unorm(w[3], 8) - 1);
    }
    return static_cast<int64_t>(N);
}

int64_t read_dvsplat(const std::vector<uint8_t>& f, float* rows, int64_t cap) {
    if (f.size() < 28) return fail(".dvsplat: no header");
    uint32_t hd[7];
    std::memcpy(hd, f.data(), 28);
    const size_t N = hd[0], chunks = hd[1];
    if (N == 0) return fail(".dvsplat: empty");
    size_t need = 28 + chunks * 24 + N * 4;
    for (int d = 0; d < 4; d++) need += static_cast<size_t>(hd[2 + d]) * static_cast<size_t>(7 + 3 * (d + 1) * (d + 1));
    if (f.size() < need || chunks * 256 < N) return fail(".dvsplat: truncated or inconsistent file");
    if (!rows) return static_cast<int64_t>(N);
    const size_t lim = std::min<size_t>(N, static_cast<size_t>(cap));
    for (size_t i = 0; i < lim; i++) {
        float* r = rows + i * kRow;
        std::fill(r, r + kRow, 0.f);
        float b[6];
        std::memcpy(b, f.data() + 28 + (i / 256) * 24, 24);
        uint32_t w;
        std::memcpy(&w, f.data() + 28 + chunks * 24 + i * 4, 4);
        unpack_in_box(w, b, b + 3, r);
    }
    size_t off = 28 + chunks * 24 + N * 4, first = 0;
    for (int d = 0; d < 4; d++) {
        const int ncoef = (d + 1) * (d + 1) - 1;
        const size_t stride = static_cast<size_t>(7 + 3 * (1 + ncoef));
        for (size_t s = 0; s < hd[2 + d] && first + s < lim; s++) {
            const uint8_t* a = f.data() + off + s * stride;
            float* r = rows + (first + s) * kRow;
            for (int k = 0; k < 3; k++) r[kRowScale + k] = static_cast<float>(a[k]) / 16.0f - 10.0f;
            float x[3];
            for (int k = 0; k < 3; k++) x[k] = static_cast<float>(a[3 + k]) * 1.0f / 127.5f + -1.0f;
            r[kRowRot] = std::sqrt(std::max(0.0f, 1.0f - (x[0] * x[0] + x[1] * x[1] + x[2] * x[2])));
            for (int k = 0; k < 3; k++) r[kRowRot + 1 + k] = x[k];
            r[kRowOpacity] = logit_of(a[6] / 255.0f);
            for (int k = 0; k < 3; k++) r[3 + k] = (a[7 + k] / 255.0f - 0.5f) / kWideColour;
            for (int j = 0; j < ncoef * 3; j++) r[6 + j] = dequant_sh(a[10 + j]);
        }
        off += hd[2 + d] * stride;
        first += hd[2 + d];
    }
    return static_cast<int64_t>(N);
}

inline float half_to_float(uint16_t h) {
    const uint32_t sign = h >> 15, e = (h >> 10) & 31, m = h & 1023;
    const float s = sign ? -1.0f : 1.0f;
    if (e == 0) return s * std::pow(2.0f, -14.0f) * static_cast<float>(m) / 1024.0f;
    if (e == 31) return m ? std::numeric_limits<float>::quiet_NaN() : s * std::numeric_limits<float>::infinity();
    return s * std::pow(2.0f, static_cast<float>(e) - 15.0f) * (1.0f + static_cast<float>(m) / 1024.0f);
}

int64_t read_spz(const std::vector<uint8_t>& gz, float* rows, int64_t cap, uint32_t* flags) {
    std::vector<uint8_t> f;
    if (!gunzip(gz, &f)) return fail(".spz: not a gzip stream");
    if (f.size() < 16) return fail(".spz: no header");
    uint32_t magic, version, count;
    std::memcpy(&magic, f.data(), 4);
    std::memcpy(&version, f.data() + 4, 4);
    std::memcpy(&count, f.data() + 8, 4);
    const int degree = f[12], frac = f[13];
    if (magic != kSpzMagic) return fail(".spz: bad magic");
    if (version < 1 || version > 3) return fail(".spz: unsupported version " + std::to_string(version));
    if (count > 10000000u || degree > 3) return fail(".spz: header out of range");
    if (f[14] & 1) *flags |= DVS_IO_ANTIALIASED;
    const size_t N = count, dim = static_cast<size_t>((degree + 1) * (degree + 1) - 1);
    const size_t pb = version == 1 ? 6 : 9, rb = version >= 3 ? 4 : 3;
    const size_t p0 = 16, a0 = p0 + pb * N, c0 = a0 + N, s0 = c0 + 3 * N, r0 = s0 + 3 * N, h0 = r0 + rb * N;
    if (f.size() < h0 + 3 * dim * N) return fail(".spz: truncated payload");
    if (!rows) return static_cast<int64_t>(N);
    const float inv = static_cast<float>(1.0 / (1 << frac));
    for (size_t i = 0; i < std::min<size_t>(N, static_cast<size_t>(cap)); i++) {
        float* r = rows + i * kRow;
        std::fill(r, r + kRow, 0.f);
        for (int a = 0; a < 3; a++) {
            if (version == 1) {
                uint16_t hv;
                std::memcpy(&hv, f.data() + p0 + 6 * i + 2 * a, 2);
                r[a] = half_to_float(hv);
            } else {
                const uint8_t* b = f.data() + p0 + 9 * i + 3 * a;
                int32_t fx = b[0] | (b[1] << 8) | (b[2] << 16);
                if (fx & 0x800000) fx |= static_cast<int32_t>(0xff000000u);
                r[a] = static_cast<float>(fx) * inv;
            }
            r[kRowScale + a] = f[s0 + 3 * i + a] / 16.0f - 10.0f;
            r[3 + a] = ((f[c0 + 3 * i + a] / 255.0f) - 0.5f) / kWideColour;
        }
        r[kRowOpacity] = logit_of(f[a0 + i] / 255.0f);
        float q[4];  // x y z w
        if (version >= 3) {
            uint32_t comp;
            std::memcpy(&comp, f.data() + r0 + 4 * i, 4);
            const int big = static_cast<int>(comp >> 30);
            float sum = 0;
            for (int k = 3; k >= 0; k--)
                if (k != big) {
                    const uint32_t mag = comp & 511u, neg = (comp >> 9) & 1u;
                    comp >>= 10;
                    q[k] = kInvSqrt2 * static_cast<float>(mag) / 511.0f;
                    if (neg) q[k] = -q[k];
                    sum += q[k] * q[k];
                }
            q[big] = std::sqrt(1.0f - sum);
        } else {
            const uint8_t* b = f.data() + r0 + 3 * i;
            for (int k = 0; k < 3; k++) q[k] = static_cast<float>(b[k]) * (1.0f / 127.5f) + -1.0f;
            q[3] = std::sqrt(std::max(0.0f, 1.0f - (q[0] * q[0] + q[1] * q[1] + q[2] * q[2])));
        }
        r[kRowRot] = q[3];
        for (int k = 0; k < 3; k++) r[kRowRot + 1 + k] = q[k];
        for (size_t j = 0; j < 3 * dim; j++) r[6 + j] = dequant_sh(f[h0 + 3 * dim * i + j]);  // [coef][rgb], zero beyond `degree`
    }
    return static_cast<int64_t>(N);
}

// tiny_gsplat.cpp:817-992 (load_reduced_ply), float and half-float positions.  The header is read the way the reference
// reads it: every `element vertex` line opens the next degree block, the line after it tells f16 from float positions, a
// one-byte property type means a codebook-quantised file (written only by callers that pass a codebook; not supported).
// Rows come back block by block (degree 0 first); coefficients a degree does not store stay 0.
int64_t read_reduced_ply(const std::vector<uint8_t>& f, float* rows, int64_t cap) {
    uint64_t count[4] = {0, 0, 0, 0};
    int blocks = 0;
    bool half = false, quantised = false;
    size_t p = 0, body = 0;
    bool after_element = false;
    while (p < f.size()) {
        const uint8_t* nl = static_cast<const uint8_t*>(std::memchr(f.data() + p, '\n', f.size() - p));
        if (!nl) break;
        const std::string line(reinterpret_cast<const char*>(f.data() + p), static_cast<size_t>(nl - (f.data() + p)));
        p = static_cast<size_t>(nl - f.data()) + 1;
        if (line == "end_header") { body = p; break; }
        if (after_element) {
            if (line.find("f16") != std::string::npos) half = true;
            after_element = false;
        }
        if (line.find("element vertex") != std::string::npos) {
            if (blocks >= 4) return fail("reduced PLY: more than four vertex blocks");
            std::istringstream ss(line);
            std::string a, b;
            ss >> a >> b >> count[blocks++];
            after_element = true;
        }
        if (line.find("property") != std::string::npos && blocks > 0) {
            std::istringstream ss(line);
            std::string a, type;
            ss >> a >> type;
            if (type == "u8") quantised = true;
            else if (type != "float" && type != "uint" && type != "f16") return fail("reduced PLY: unrecognized type " + type);
        }
    }
    if (!body) return fail("reduced PLY: no end_header");
    if (blocks != 4) return fail("reduced PLY: expected four vertex blocks (one per SH degree)");
    if (quantised) return fail("reduced PLY: codebook-quantised files are not supported");
    const int64_t n = static_cast<int64_t>(count[0] + count[1] + count[2] + count[3]);
    if (n <= 0) return fail("reduced PLY: no vertices");
    const size_t xyz = half ? 6 : 12;
    size_t need = 0;
    for (int d = 0; d < 4; d++) {
        const size_t stride = xyz + 12 + 12 * static_cast<size_t>((d + 1) * (d + 1) - 1) + 4 + 12 + 16;
        if (count[d] > (f.size() - body) / stride) return fail("reduced PLY: truncated payload");  // also keeps `need` from wrapping
        need += static_cast<size_t>(count[d]) * stride;
    }
    if (f.size() - body < need) return fail("reduced PLY: truncated payload");
    if (!rows) return n;
    const uint8_t* src = f.data() + body;
    int64_t i = 0;
    for (int d = 0; d < 4; d++) {
        const int coeffs = (d + 1) * (d + 1) - 1;
        const size_t stride = xyz + 12 + 12 * static_cast<size_t>(coeffs) + 4 + 12 + 16;
        for (uint64_t k = 0; k < count[d]; k++, i++, src += stride) {
            if (i >= cap) continue;
            float* r = rows + i * kRow;
            for (int a = 0; a < kRow; a++) r[a] = 0.f;
            if (half) {
                for (int a = 0; a < 3; a++) {
                    uint16_t hv;
                    std::memcpy(&hv, src + 2 * a, 2);
                    r[a] = half_to_float(hv);
                }
            } else {
                std::memcpy(r, src, 12);
            }
            std::memcpy(r + 3, src + xyz, 12 + 12 * static_cast<size_t>(coeffs));  // f_dc, then the stored coefficients in order
            const uint8_t* t = src + xyz + 12 + 12 * static_cast<size_t>(coeffs);
            std::memcpy(r + kRowOpacity, t, 4);
            std::memcpy(r + kRowScale, t + 4, 12);
            std::memcpy(r + kRowRot, t + 16, 16);
        }
    }
    return n;
}

bool ends_with(const std::string& s, const char* suf) {
    const size_t n = std::strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

}  // namespace

extern "C" {

DVS_API int dvs_model_format_from_path(const char* path) {
    if (!path) return 0;
    const std::string p(path);
    if (ends_with(p, ".ply")) {  // gaussian_model.cpp:439-451: ".compressed" is looked for first, then ".reduced"
        if (p.find(".compressed") != std::string::npos) return DVS_FMT_COMPRESSED_PLY;
        return p.find(".reduced") != std::string::npos ? DVS_FMT_REDUCED_PLY : DVS_FMT_PLY;
    }
    if (ends_with(p, ".splat")) return DVS_FMT_SPLAT;
    if (ends_with(p, ".dvsplat")) return DVS_FMT_DVSPLAT;
    if (ends_with(p, ".spz")) return DVS_FMT_SPZ;
    return 0;
}

DVS_API int dvs_model_write(const char* path, int format, int64_t N, const float* means3D, const float* sh0,
                            const float* shN, const float* logit_opac, const float* log_scales, const float* quats,
                            const uint8_t* degrees, uint32_t flags) {
    g_err.clear();
    if (!path || N < 0 || (N > 0 && !(means3D && sh0 && shN && logit_opac && log_scales && quats)))
        return fail("dvs_model_write: null argument");
    if (format == DVS_FMT_AUTO) format = dvs_model_format_from_path(path);
    const Cloud c{N, means3D, sh0, shN, logit_opac, log_scales, quats, degrees};
    if (N == 0 && format != DVS_FMT_PLY) return fail("dvs_model_write: an empty model can only be written as PLY");
    switch (format) {
        case DVS_FMT_PLY: return write_ply(path, c, flags);
        case DVS_FMT_SPLAT: return write_splat(path, c);
        case DVS_FMT_COMPRESSED_PLY: return write_compressed_ply(path, c, flags);
        case DVS_FMT_DVSPLAT: return write_dvsplat(path, c);
        case DVS_FMT_SPZ: return write_spz(path, c, flags);
        case DVS_FMT_REDUCED_PLY: return write_reduced_ply(path, c, flags);
        default: return fail(std::string("dvs_model_write: unknown model format for ") + path);
    }
}

DVS_API int64_t dvs_model_read(const char* path, int format, float* rows, int64_t cap, uint32_t* flags_out) {
    g_err.clear();
    if (!path) return fail("dvs_model_read: null path");
    if (format == DVS_FMT_AUTO) format = dvs_model_format_from_path(path);
    std::vector<uint8_t> f;
    if (!slurp(path, &f)) return fail(std::string("cannot read ") + path);
    uint32_t flags = 0;
    if (!rows) cap = 0;
    int64_t n;
    switch (format) {
        case DVS_FMT_PLY: n = read_ply(f, rows, cap, &flags); break;
        case DVS_FMT_SPLAT: n = read_splat(f, rows, cap); break;
        case DVS_FMT_COMPRESSED_PLY: n = read_compressed_ply(f, rows, cap, &flags); break;
        case DVS_FMT_DVSPLAT: n = read_dvsplat(f, rows, cap); break;
        case DVS_FMT_SPZ: n = read_spz(f, rows, cap, &flags); break;
        case DVS_FMT_REDUCED_PLY: n = read_reduced_ply(f, rows, cap); break;
        default: return fail(std::string("dvs_model_read: unknown model format for ") + path);
    }
    if (flags_out) *flags_out = flags;
    return n;
}

DVS_API const char* dvs_model_io_last_error(void) { return g_err.c_str(); }
}
