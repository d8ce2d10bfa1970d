// png.cu -- the 01 -> 02 wire format written on the device (sm_100a).
//
// The reference PNG-encodes every binarized frame with cv2.imencode on the host (12.7 ms per 1080p frame,
// R/AccessMath/preprocessing/video_worker/FCN_lecturenet_binarizer.py:56) and stage 02 reads it back with
// cv2.imdecode(raw, IMREAD_GRAYSCALE) (R/AccessMath/preprocessing/content/helper.py:31).  am_png1_encode writes, straight from the
// bit-packed ink mask the FCN epilogue produced, a PNG every decoder expands to the same 0 / 255 pixels:
//   1-bit grayscale (ink = 1 = white), filter type 0, one IDAT chunk holding a zlib stream of "stored" deflate blocks.
// Two launches per batch: scanline bytes are bit-reversed mask bytes (PNG packs the leftmost pixel into the MSB), Adler-32 is a
// block reduction of two 64-bit sums, CRC-32 is computed per 1024th of the chunk by table lookup and the partial CRCs are
// merged by a shared-memory tree with precomputed GF(2) shift matrices (the crc32_combine construction).  HBM-bound: reads
// W*H/8 bytes, writes W*H/8 + H + 77 bytes per frame (260 KB at 1080p, against 40-450 KB for cv2's deflate).
#include "am_common.cuh"
#include "../../include/accessmath_b200.h"
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace {

constexpr int kThreads = 1024;
constexpr int kLevels = 10;                      // log2(kThreads)

struct PngPlan {                                 // everything that depends only on (width, height)
    int W, H, WPR, row_bytes;                    // row_bytes = 1 (filter) + ceil(W / 8)
    long long raw, n_blocks, zlen, total;        // uncompressed bytes, stored blocks, zlib stream bytes, file bytes
    int seg;                                     // bytes per CRC segment (per thread)
    uint8_t head[41];                            // signature + IHDR chunk + IDAT length + "IDAT"
    uint32_t* d_tables;                          // [256] CRC table, then kLevels x [32] shift matrices for seg * 2^level bytes
};

uint32_t crc_table_entry(uint32_t n) {
    for (int k = 0; k < 8; ++k) n = (n & 1) ? 0xEDB88320u ^ (n >> 1) : n >> 1;
    return n;
}
uint32_t crc_bytes(const uint8_t* p, size_t n) {
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) c = crc_table_entry((c ^ p[i]) & 0xFF) ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}
uint32_t gf2_times(const uint32_t* mat, uint32_t vec) {
    uint32_t sum = 0;
    for (int i = 0; vec; vec >>= 1, ++i) if (vec & 1) sum ^= mat[i];
    return sum;
}
void gf2_square(uint32_t* sq, const uint32_t* mat) { for (int i = 0; i < 32; ++i) sq[i] = gf2_times(mat, mat[i]); }
void gf2_mul(uint32_t* out, const uint32_t* a, const uint32_t* b) { for (int i = 0; i < 32; ++i) out[i] = gf2_times(a, b[i]); }   // a after b
// operator that advances a CRC register over `nbytes` zero bytes
void crc_shift_matrix(long long nbytes, uint32_t* out) {
    uint32_t bit[32], tmp[32], pw[32];
    bit[0] = 0xEDB88320u;                                            // one zero BIT
    for (int i = 1; i < 32; ++i) bit[i] = 1u << (i - 1);
    gf2_square(tmp, bit); gf2_square(bit, tmp); gf2_square(pw, bit); // 2, 4, 8 bits = one byte
    for (int i = 0; i < 32; ++i) out[i] = 1u << i;                   // identity
    while (nbytes) {
        if (nbytes & 1) { gf2_mul(tmp, pw, out); memcpy(out, tmp, sizeof(tmp)); }
        gf2_square(tmp, pw); memcpy(pw, tmp, sizeof(tmp));
        nbytes >>= 1;
    }
}
void be32(uint8_t* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

std::mutex g_mu;
std::map<std::tuple<int, int, int, int>, PngPlan*> g_plans;

constexpr int kLaneBytes = 256;                  // deflate writer: raw bytes one lane tokenises (a run never exceeds the 258-byte match)
constexpr int kSegBytes = 32 * kLaneBytes;       // ... and one warp turns into one fixed-Huffman block
long long deflate_capacity(long long raw) {      // zlib stream bytes in the worst case: every byte a 9-bit literal
    const long long n_seg = (raw + kSegBytes - 1) / kSegBytes;
    return 2 + (raw * 9 + 7) / 8 + 8 * n_seg + 4;
}

// deflate = 0: the stored-block container (fixed size); 1: the fixed-Huffman writer, zlen / total are then CAPACITIES; 2: the same
// writer for 8-bit grayscale frames (one byte per pixel)
int png_plan(int width, int height, int deflate, PngPlan** out) {
    int dev = 0;
    AM_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_tuple(dev, width, height, deflate);
    auto it = g_plans.find(key);
    if (it == g_plans.end()) {
        PngPlan* p = new PngPlan();
        p->W = width; p->H = height; p->WPR = am_words_per_row_impl(width);
        p->row_bytes = 1 + (deflate == 2 ? width : (width + 7) / 8);
        p->raw = (long long)height * p->row_bytes;
        p->n_blocks = (p->raw + 65534) / 65535;
        p->zlen = deflate ? deflate_capacity(p->raw) : 2 + p->raw + 5 * p->n_blocks + 4;
        p->total = 8 + 25 + 12 + p->zlen + 12;
        if (deflate) p->total = (p->total + 15) & ~15LL;                // frames start on word boundaries (bit writer: 32-bit atomicOr)
        const long long crc_len = 4 + p->zlen;                       // "IDAT" + chunk data
        p->seg = (int)((crc_len + kThreads - 1) / kThreads);
        uint8_t* h = p->head;
        const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
        memcpy(h, sig, 8);
        be32(h + 8, 13); memcpy(h + 12, "IHDR", 4); be32(h + 16, (uint32_t)width); be32(h + 20, (uint32_t)height);
        h[24] = deflate == 2 ? 8 : 1; h[25] = 0; h[26] = 0; h[27] = 0; h[28] = 0;   // bit depth, grayscale, deflate, filter method 0, no interlace
        be32(h + 29, crc_bytes(h + 12, 17));
        be32(h + 33, (uint32_t)p->zlen); memcpy(h + 37, "IDAT", 4);
        std::vector<uint32_t> t(256 + kLevels * 32);
        for (uint32_t n = 0; n < 256; ++n) t[n] = crc_table_entry(n);
        for (int l = 0; l < kLevels; ++l) crc_shift_matrix((long long)p->seg << l, &t[256 + 32 * l]);
        AM_CUDA(cudaMalloc(&p->d_tables, t.size() * 4));
        AM_CUDA(cudaMemcpy(p->d_tables, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
        it = g_plans.emplace(key, p).first;
    }
    *out = it->second;
    return AM_OK;
}

struct PngArgs {
    int W, H, WPR, row_bytes, seg;
    int depth;                                   // 1: source = bit-packed masks; 8: source = uint8 frames [H][W]
    long long raw, n_blocks, zlen, total;
    uint8_t head[41];
};

__device__ __forceinline__ uint32_t gf2_times_dev(const uint32_t* __restrict__ mat, uint32_t vec) {
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) sum ^= ((vec >> i) & 1u) ? mat[i] : 0u;
    return sum;
}

constexpr int kSliceBlocks = 16;                 // CTAs per frame in the scanline kernel

// Kernel A, grid (kSliceBlocks, frames): file header, stored-block headers, scanline bytes of this CTA's slice of the raw stream and
// its Adler-32 partial sums  (A = 1 + sum d_i,  B = n + sum (n - i) d_i  mod 65521)  ->  partial[frame][slice] = (sum d, sum (n - i) d)
__global__ void __launch_bounds__(kThreads)
k_png1_scanlines(const uint32_t* __restrict__ bits, const PngArgs a, uint8_t* __restrict__ out, unsigned long long* __restrict__ partial) {
    __shared__ unsigned long long s_red[2][32];
    const int f = blockIdx.y, tid = threadIdx.x;
    const uint32_t* fb = bits + (size_t)f * a.H * a.WPR;
    uint8_t* o = out + (size_t)f * a.total;
    uint8_t* z = o + 41;                                              // zlib stream
    if (blockIdx.x == 0) {
        if (tid < 41) o[tid] = a.head[tid];
        if (tid == 0) { z[0] = 0x78; z[1] = 0x01; }
        for (long long b = tid; b < a.n_blocks; b += kThreads) {      // stored-block headers: BFINAL, LEN, ~LEN (little endian)
            const long long start = b * 65535, len = min(65535LL, a.raw - start);
            uint8_t* h = z + 2 + start + 5 * b;
            h[0] = (b == a.n_blocks - 1) ? 1 : 0;
            h[1] = (uint8_t)(len & 0xFF); h[2] = (uint8_t)(len >> 8); h[3] = (uint8_t)(~len & 0xFF); h[4] = (uint8_t)((~len >> 8) & 0xFF);
        }
    }
    unsigned long long sa = 0, sb = 0;
    const int last_bits = a.W & 7;
    const unsigned raw = (unsigned)a.raw, rb = (unsigned)a.row_bytes;  // raw < 2^31 (checked on the host): 32-bit index arithmetic
    const unsigned per = (raw + gridDim.x - 1) / gridDim.x, r0 = blockIdx.x * per, r1 = min(raw, r0 + per);
    for (unsigned r = r0 + tid; r < r1; r += kThreads) {
        const unsigned y = r / rb, c = r - y * rb;
        uint32_t d = 0;
        if (c > 0) {
            const unsigned k = c - 1;                                 // byte k of the row = pixels 8k .. 8k+7
            d = (fb[(size_t)y * a.WPR + (k >> 2)] >> (8 * (k & 3))) & 0xFFu;
            if (last_bits && k == rb - 2) d &= (1u << last_bits) - 1u;
            d = __brev(d) >> 24;                                      // leftmost pixel into the most significant bit
        }
        z[2 + r + 5 * (r / 65535u + 1)] = (uint8_t)d;
        sa += d;
        sb += (unsigned long long)(raw - r) * d;
    }
    for (int off = 16; off; off >>= 1) {                              // 64-bit warp reductions
        sa += __shfl_down_sync(0xffffffffu, sa, off);
        sb += __shfl_down_sync(0xffffffffu, sb, off);
    }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = sa; s_red[1][tid >> 5] = sb; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long A = 0, B = 0;
        for (int w = 0; w < kThreads / 32; ++w) { A += s_red[0][w]; B += s_red[1][w]; }
        partial[2 * ((size_t)f * gridDim.x + blockIdx.x)] = A;
        partial[2 * ((size_t)f * gridDim.x + blockIdx.x) + 1] = B;
    }
}

// Kernel B, one CTA per frame: Adler-32 from the partial sums, then CRC-32 of "IDAT" + data (thread e, counted from the END of
// the region, owns bytes [end - (e+1) seg, end - e seg); partial CRCs merged by a tree of zero-shift matrices), then IEND.
// d_zlen != NULL (deflate writer): the zlib stream of frame f is d_zlen[f] bytes long (<= a.zlen, which then is the CAPACITY the CRC
// segment size was derived from); the IDAT length field is patched and the file size goes to d_sizes[f].
__global__ void __launch_bounds__(kThreads)
k_png1_checksums(PngArgs a, const uint32_t* __restrict__ tables, const unsigned long long* __restrict__ partial, int n_slices,
                 uint8_t* __restrict__ out, const long long* __restrict__ d_zlen, long long* __restrict__ d_sizes) {
    __shared__ uint32_t s_tab[256];
    __shared__ uint32_t s_mat[kLevels * 32];
    __shared__ uint32_t s_crc[kThreads];
    const int f = blockIdx.x, tid = threadIdx.x;
    uint8_t* o = out + (size_t)f * a.total;
    uint8_t* z = o + 41;
    if (d_zlen) {
        a.zlen = d_zlen[f];
        if (tid == 0) {
            o[33] = (uint8_t)(a.zlen >> 24); o[34] = (uint8_t)(a.zlen >> 16); o[35] = (uint8_t)(a.zlen >> 8); o[36] = (uint8_t)a.zlen;
            d_sizes[f] = 41 + a.zlen + 16;
        }
    }
    if (tid < 256) s_tab[tid] = tables[tid];
    if (tid < kLevels * 32) s_mat[tid] = tables[256 + tid];
    if (tid == 0 && !d_zlen) {                                        // (the deflate writer has placed its Adler-32 itself)
        unsigned long long A = 1, B = (unsigned long long)(a.raw % 65521);
        for (int w = 0; w < n_slices; ++w) { A += partial[2 * ((size_t)f * n_slices + w)]; B = (B + partial[2 * ((size_t)f * n_slices + w) + 1] % 65521) % 65521; }
        const uint32_t adler = (uint32_t)((B % 65521) << 16) | (uint32_t)(A % 65521);
        uint8_t* p = z + a.zlen - 4;
        p[0] = adler >> 24; p[1] = adler >> 16; p[2] = adler >> 8; p[3] = adler;
    }
    __syncthreads();                                                  // the whole chunk is in place (block-scope visibility)
    const long long crc_len = 4 + a.zlen;
    const uint8_t* reg = o + 37;                                      // "IDAT"
    {
        const int e = tid;
        long long hi = crc_len - (long long)e * a.seg, lo = hi - a.seg;
        uint32_t c = 0;                                               // CRC of an empty string
        if (hi > 0) {
            if (lo < 0) lo = 0;
            c = 0xFFFFFFFFu;
            long long i = lo;
            for (; i + 8 <= hi; i += 8) {                             // the eight loads do not depend on the CRC chain
                uint32_t b[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) b[j] = reg[i + j];
#pragma unroll
                for (int j = 0; j < 8; ++j) c = s_tab[(c ^ b[j]) & 0xFFu] ^ (c >> 8);
            }
            for (; i < hi; ++i) c = s_tab[(c ^ reg[i]) & 0xFFu] ^ (c >> 8);
            c ^= 0xFFFFFFFFu;
        }
        s_crc[e] = c;
    }
    __syncthreads();
    // tree: node (e, level) = crc of 2^level segments ending at segment e; crc(A || B) = shift(crc A, |B|) ^ crc B
    for (int l = 0; l < kLevels; ++l) {
        const int stride = 1 << l;
        uint32_t merged = 0;
        const bool act = (tid & (2 * stride - 1)) == 0;
        if (act) {
            const uint32_t right = s_crc[tid], left = s_crc[tid + stride];      // `left` lies earlier in the stream
            const long long left_hi = crc_len - (long long)(tid + stride) * a.seg;
            merged = (left_hi > 0) ? (gf2_times_dev(s_mat + 32 * l, left) ^ right) : right;
        }
        __syncthreads();
        if (act) s_crc[tid] = merged;
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t crc = s_crc[0];
        uint8_t* p = z + a.zlen;
        p[0] = crc >> 24; p[1] = crc >> 16; p[2] = crc >> 8; p[3] = crc;
        const uint8_t iend[12] = {0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xAE, 0x42, 0x60, 0x82};
        for (int i = 0; i < 12; ++i) p[4 + i] = iend[i];
    }
}


// ------------------------------------------------------------------------------------------------------------------------------
// Deflate writer: the same 1-bit PNG with a COMPRESSED zlib stream (RFC 1951 fixed Huffman codes, run-length matches of distance 1).
// Whiteboard / chalkboard masks are > 95 % zero bytes: a run of n equal bytes becomes one literal + one (length n - 1, distance 1)
// match, 26 bits per 256 empty bytes.  No serial pass over the stream:
//   * warp-segment s = raw bytes [s * 8192, (s + 1) * 8192) is ONE deflate block; lane l tokenises bytes [256 l, 256 l + 256) of it
//     on its own (runs are cut at lane boundaries), pass 1 counts bits, a warp scan + a CTA scan over the segments place every lane;
//   * blocks end on a byte boundary -- a fixed block is followed by an EMPTY stored block (3 header bits, pad, 00 00 FF FF), the
//     classic sync-flush marker -- so segment offsets are byte counts and the stream stays one valid deflate stream;
//   * the emit pass re-tokenises and ORs the codes into the zero-initialised output with 32-bit atomics (neighbouring lanes share words).
// Adler-32 over the raw bytes falls out of the counting pass; k_png1_checksums (above) adds the CRC-32 over the variable-length chunk.
struct BitWriter {
    uint32_t* words; unsigned long long w; uint64_t acc; int fill;
    __device__ void init(uint32_t* base, unsigned long long bitpos) { words = base; w = bitpos >> 5; fill = (int)(bitpos & 31); acc = 0; }
    __device__ void put(uint32_t code, int n) {                     // n <= 32 bits, LSB first
        acc |= (uint64_t)code << fill;
        fill += n;
        if (fill >= 32) { atomicOr(&words[w], (uint32_t)acc); ++w; acc >>= 32; fill -= 32; }
    }
    __device__ void flush() { if (fill > 0) atomicOr(&words[w], (uint32_t)acc); fill = 0; acc = 0; }
    __device__ unsigned long long bitpos() const { return (w << 5) + (unsigned)fill; }
};
// literal / length symbol -> (code bits reversed for the LSB-first stream, length); RFC 1951 3.2.6
__device__ __forceinline__ uint32_t fixed_code(uint32_t sym, int& n) {
    uint32_t code;
    if (sym < 144) { code = 0x30 + sym; n = 8; }
    else if (sym < 256) { code = 0x190 + (sym - 144); n = 9; }
    else if (sym < 280) { code = sym - 256; n = 7; }
    else { code = 0xC0 + (sym - 280); n = 8; }
    return __brev(code) >> (32 - n);
}
// match length 3 .. 258 -> (symbol, extra-bit count, extra value); RFC 1951 3.2.5
__device__ __forceinline__ void length_symbol(int len, uint32_t& sym, int& eb, uint32_t& ev) {
    if (len == 258) { sym = 285; eb = 0; ev = 0; return; }
    const int l = len - 3;
    if (l < 8) { sym = 257 + l; eb = 0; ev = 0; return; }
    eb = (31 - __clz(l)) - 2;
    sym = 265 + 4 * (eb - 1) + ((l >> eb) - 4);
    ev = (uint32_t)l & ((1u << eb) - 1u);
}
// A warp's segment of the raw stream, staged in shared memory: byte i of the segment at sm[(i / 256) * 260 + i % 256] (every lane's
// 256 bytes start 65 words apart: conflict-free when all lanes walk their ranges in step).  Raw stream byte r of a frame = the filter
// byte (0) in front of every scanline, else 8 pixels with the leftmost in the MSB.  Coalesced, independent loads: a lane that fetched
// its bytes one by one from global memory was latency bound at ~1 us per byte (240 us per pass and batch, profile r02_h).
constexpr int kLanePitch = kLaneBytes + 4;
__device__ __forceinline__ void stage_segment(const uint32_t* __restrict__ fb, const PngArgs& a, unsigned seg_r0, uint8_t* sm, int lane) {
    const unsigned raw = (unsigned)a.raw, rb = (unsigned)a.row_bytes, last_bits = (unsigned)a.W & 7u;
    if (a.depth == 8) {                                              // 8-bit grayscale: the scanline bytes are the pixels themselves
        const uint8_t* px = (const uint8_t*)fb;
#pragma unroll 4
        for (unsigned i = lane; i < (unsigned)kSegBytes; i += 32) {
            const unsigned r = seg_r0 + i;
            uint32_t d = 0;
            if (r < raw) {
                const unsigned y = r / rb, c = r - y * rb;
                if (c > 0) d = px[(size_t)y * a.W + (c - 1)];
            }
            sm[(i >> 8) * kLanePitch + (i & 255u)] = (uint8_t)d;
        }
        __syncwarp();
        return;
    }
#pragma unroll 4
    for (unsigned i = lane; i < (unsigned)kSegBytes; i += 32) {
        const unsigned r = seg_r0 + i;
        uint32_t d = 0;
        if (r < raw) {
            const unsigned y = r / rb, c = r - y * rb;
            if (c > 0) {
                const unsigned k = c - 1;
                d = (fb[(size_t)y * a.WPR + (k >> 2)] >> (8 * (k & 3))) & 0xFFu;
                if (last_bits && k == rb - 2) d &= (1u << last_bits) - 1u;
                d = __brev(d) >> 24;
            }
        }
        sm[(i >> 8) * kLanePitch + (i & 255u)] = (uint8_t)d;
    }
    __syncwarp();
}
// tokens of the n raw bytes b[0 .. n) (stream position r0): bit count (kEmit = false) or emission; also the lane's Adler partial sums
// in the counting pass
template <bool kEmit>
__device__ __forceinline__ unsigned tokenize(const uint8_t* b, unsigned n, unsigned r0, unsigned raw, BitWriter* bw,
                                             unsigned long long* sa, unsigned long long* sb) {
    unsigned nbits = 0, i = 0;
    while (i < n) {
        const uint32_t cur = b[i];
        unsigned run = 1;
        while (i + run < n && b[i + run] == cur) ++run;
        if (!kEmit) {
            const unsigned r = r0 + i;
            *sa += (unsigned long long)cur * run;
            *sb += (unsigned long long)cur * ((unsigned long long)run * (raw - r) - (unsigned long long)run * (run - 1) / 2);
        }
        int ln; const uint32_t lc = fixed_code(cur, ln);
        if (run >= 4) {                                             // literal + (length run - 1, distance 1)
            uint32_t sym, ev; int eb, sn;
            length_symbol((int)run - 1, sym, eb, ev);
            const uint32_t sc = fixed_code(sym, sn);
            if (kEmit) { bw->put(lc, ln); bw->put(sc | (ev << sn), sn + eb); bw->put(0, 5); }
            nbits += ln + sn + eb + 5;
        } else {
            if (kEmit) for (unsigned k = 0; k < run; ++k) bw->put(lc, ln);
            nbits += ln * run;
        }
        i += run;
    }
    return nbits;
}

// Three launches per batch, a warp per segment (the only dependency between segments is the scan of their byte sizes):
constexpr int kDefWarps = 4;                     // warps (= segments) per CTA of the count / emit kernels

// pass 1, grid (ceil(n_seg / kDefWarps), frames): bits per lane -> lane_off[frame][seg][lane]; bytes per segment; Adler partial sums
__global__ void __launch_bounds__(kDefWarps * 32)
k_png1_deflate_count(const uint32_t* __restrict__ bits, const PngArgs a, int n_seg, unsigned* __restrict__ lane_off,
                     unsigned* __restrict__ seg_bytes, unsigned long long* __restrict__ seg_adler) {
    const int f = blockIdx.y, lane = threadIdx.x & 31, s = blockIdx.x * kDefWarps + (threadIdx.x >> 5);
    if (s >= n_seg) return;
    __shared__ __align__(16) uint8_t s_seg[kDefWarps][32 * kLanePitch];
    const uint32_t* fb = a.depth == 8 ? (const uint32_t*)((const uint8_t*)bits + (size_t)f * a.H * a.W) : bits + (size_t)f * a.H * a.WPR;
    const unsigned raw = (unsigned)a.raw;
    const unsigned r0 = min(raw, (unsigned)s * kSegBytes + (unsigned)lane * kLaneBytes), r1 = min(raw, r0 + kLaneBytes);
    uint8_t* sm = s_seg[threadIdx.x >> 5];
    stage_segment(fb, a, (unsigned)s * kSegBytes, sm, lane);
    unsigned long long sa = 0, sb = 0;
    const unsigned nb = tokenize<false>(sm + lane * kLanePitch, r1 - r0, r0, raw, nullptr, &sa, &sb);
    unsigned incl = nb;
    for (int d = 1; d < 32; d <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    lane_off[((size_t)f * n_seg + s) * 32 + lane] = 3 + incl - nb;   // behind the 3-bit block header
    for (int off = 16; off; off >>= 1) { sa += __shfl_down_sync(0xffffffffu, sa, off); sb += __shfl_down_sync(0xffffffffu, sb, off); }
    if (lane == 31) {
        const unsigned body = 3 + incl + 7;                          // header, tokens, end-of-block
        seg_bytes[(size_t)f * n_seg + s] = (s == n_seg - 1) ? (body + 7) / 8 : (body + 3 + 7) / 8 + 4;   // (the last block is followed by Adler-32)
    }
    if (lane == 0) { seg_adler[2 * ((size_t)f * n_seg + s)] = sa; seg_adler[2 * ((size_t)f * n_seg + s) + 1] = sb; }
}

// pass 2, one warp per frame: exclusive scan of the segment sizes (in place), Adler-32, stream length; file header
__global__ void k_png1_deflate_scan(const PngArgs a, int n_seg, unsigned* __restrict__ seg_bytes, const unsigned long long* __restrict__ seg_adler,
                                    uint32_t* __restrict__ adler_out, long long* __restrict__ d_zlen, uint8_t* __restrict__ out) {
    const int f = blockIdx.x, lane = threadIdx.x;
    uint8_t* o = out + (size_t)f * a.total;
    for (int i = lane; i < 41; i += 32) o[i] = a.head[i];             // (IDAT length is patched by k_png1_checksums)
    if (lane == 0) { o[41] = 0x78; o[42] = 0x01; }
    unsigned* sz = seg_bytes + (size_t)f * n_seg;
    unsigned carry = 0;
    unsigned long long A = 0, B = 0;
    for (int base = 0; base < n_seg; base += 32) {
        const bool ok = base + lane < n_seg;
        const unsigned v = ok ? sz[base + lane] : 0;
        unsigned incl = v;
        for (int d = 1; d < 32; d <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        if (ok) {
            sz[base + lane] = carry + incl - v;
            A += seg_adler[2 * ((size_t)f * n_seg + base + lane)];
            B = (B + seg_adler[2 * ((size_t)f * n_seg + base + lane) + 1] % 65521) % 65521;
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    for (int off = 16; off; off >>= 1) { A += __shfl_down_sync(0xffffffffu, A, off); B += __shfl_down_sync(0xffffffffu, B, off); }
    if (lane == 0) {
        A += 1; B += (unsigned long long)(a.raw % 65521);
        adler_out[f] = (uint32_t)((B % 65521) << 16) | (uint32_t)(A % 65521);
        d_zlen[f] = 2 + (long long)carry + 4;                        // zlib header + blocks + Adler-32
    }
}

// pass 3, same grid as pass 1: emit
__global__ void __launch_bounds__(kDefWarps * 32)
k_png1_deflate_emit(const uint32_t* __restrict__ bits, const PngArgs a, int n_seg, const unsigned* __restrict__ lane_off,
                    const unsigned* __restrict__ seg_off, const uint32_t* __restrict__ adler, uint8_t* __restrict__ out) {
    const int f = blockIdx.y, lane = threadIdx.x & 31, s = blockIdx.x * kDefWarps + (threadIdx.x >> 5);
    if (s >= n_seg) return;
    __shared__ __align__(16) uint8_t s_seg[kDefWarps][32 * kLanePitch];
    const uint32_t* fb = a.depth == 8 ? (const uint32_t*)((const uint8_t*)bits + (size_t)f * a.H * a.W) : bits + (size_t)f * a.H * a.WPR;
    uint32_t* words = (uint32_t*)(out + (size_t)f * a.total);         // frame buffers start on 16-byte boundaries
    const unsigned raw = (unsigned)a.raw;
    const unsigned r0 = min(raw, (unsigned)s * kSegBytes + (unsigned)lane * kLaneBytes), r1 = min(raw, r0 + kLaneBytes);
    uint8_t* sm = s_seg[threadIdx.x >> 5];
    stage_segment(fb, a, (unsigned)s * kSegBytes, sm, lane);
    const unsigned long long seg_bit = 8ull * (43 + seg_off[(size_t)f * n_seg + s]);
    const bool last = s == n_seg - 1;
    BitWriter bw;
    if (lane == 0) { bw.init(words, seg_bit); bw.put(last ? 3u : 2u, 3); }               // BFINAL, BTYPE = 01 (fixed Huffman), LSB first
    else bw.init(words, seg_bit + lane_off[((size_t)f * n_seg + s) * 32 + lane]);
    tokenize<true>(sm + lane * kLanePitch, r1 - r0, r0, raw, &bw, nullptr, nullptr);
    if (lane == 31) {
        bw.put(0, 7);                                                // end of block (symbol 256)
        if (!last) bw.put(0, 3);                                     // empty stored block: BFINAL = 0, BTYPE = 00 ...
        const int pad = (int)((8 - (bw.bitpos() & 7)) & 7);
        if (pad) bw.put(0, pad);                                     // ... pad to the byte boundary ...
        if (!last) { bw.put(0x0000u, 16); bw.put(0xFFFFu, 16); }     // ... LEN = 0, NLEN = 0xFFFF
        else bw.put(__byte_perm(adler[f], 0, 0x0123), 32);           // Adler-32, big endian
    }
    bw.flush();
}

// PNG scanlines (what zlib.decompress returns for a 1-bit, filter-0 PNG: per row one filter byte + ceil(W / 8) bytes, leftmost pixel in
// the MSB) -> device-layout mask words; one thread per word.  The decode half of the wire format that has to run on the device.
__global__ void k_png1_scan_to_bits(const uint8_t* __restrict__ scan, int H, int W, int WPR, int rb, uint32_t* __restrict__ bits) {
    const int f = blockIdx.z, y = blockIdx.y, w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= WPR) return;
    const uint8_t* row = scan + ((size_t)f * H + y) * (size_t)(1 + rb) + 1;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int b = 4 * w + k;
        uint32_t d = b < rb ? row[b] : 0u;
        d = __brev(d) >> 24;
        if (b == rb - 1 && (W & 7)) d &= (1u << (W & 7)) - 1u;
        v |= d << (8 * k);
    }
    bits[((size_t)f * H + y) * WPR + w] = v;
}

}  // namespace

// scratch for the Adler partial sums (kSliceBlocks x 2 per frame; the deflate writer keeps its per-frame stream length in the same
// block): cached per device, grown on demand (stream-ordered use only)
static int png_scratch(int batch, unsigned long long** out) {
    static std::mutex mu;
    static std::map<int, std::pair<unsigned long long*, int>> scratch;
    int dev = 0;
    AM_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    auto& sc = scratch[dev];
    if (sc.second < batch) {
        if (sc.first) cudaFree(sc.first);
        sc.first = nullptr; sc.second = 0;
        AM_CUDA(cudaMalloc(&sc.first, (size_t)batch * (kSliceBlocks * 2 + 2) * sizeof(unsigned long long)));
        sc.second = batch;
    }
    *out = sc.first;
    return AM_OK;
}

extern "C" long long am_png1_size(int width, int height) {
    if (width <= 0 || height <= 0) return 0;
    const long long raw = (long long)height * (1 + (width + 7) / 8);
    return 8 + 25 + 12 + (2 + raw + 5 * ((raw + 65534) / 65535) + 4) + 12;
}

extern "C" int am_png1_encode(const uint32_t* d_bits, int batch, int height, int width, uint8_t* d_out, void* stream) {
    if (!d_bits || !d_out || batch <= 0 || width <= 0 || height <= 0) return AM_ERR_ARG;
    PngPlan* p = nullptr;
    int rc = png_plan(width, height, 0, &p);
    if (rc) return rc;
    if (p->zlen > 0x7FFFFFFFLL) return AM_ERR_ARG;                    // chunk length field; also keeps raw < 2^31
    PngArgs a;
    a.W = p->W; a.H = p->H; a.WPR = p->WPR; a.row_bytes = p->row_bytes; a.seg = p->seg; a.depth = 1;
    a.raw = p->raw; a.n_blocks = p->n_blocks; a.zlen = p->zlen; a.total = p->total;
    memcpy(a.head, p->head, 41);
    unsigned long long* d_partial = nullptr;
    rc = png_scratch(batch, &d_partial);
    if (rc) return rc;
    k_png1_scanlines<<<dim3(kSliceBlocks, batch), kThreads, 0, (cudaStream_t)stream>>>(d_bits, a, d_out, d_partial);
    k_png1_checksums<<<batch, kThreads, 0, (cudaStream_t)stream>>>(a, p->d_tables, d_partial, kSliceBlocks, d_out, nullptr, nullptr);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}

// ---- compressed variant ---------------------------------------------------------------------------------------------------------
extern "C" long long am_png1_capacity(int width, int height) {
    if (width <= 0 || height <= 0) return 0;
    const long long raw = (long long)height * (1 + (width + 7) / 8);
    return ((8 + 25 + 12 + deflate_capacity(raw) + 12) + 15) & ~15LL;
}

static int png_encode_deflate(const uint32_t* d_bits, int depth, int batch, int height, int width, uint8_t* d_out, long long* d_sizes,
                              void* stream) {
    if (!d_bits || !d_out || !d_sizes || batch <= 0 || width <= 0 || height <= 0) return AM_ERR_ARG;
    if (((uintptr_t)d_out & 15) != 0) return AM_ERR_ARG;
    PngPlan* p = nullptr;
    int rc = png_plan(width, height, depth == 8 ? 2 : 1, &p);
    if (rc) return rc;
    const long long n_seg = (p->raw + kSegBytes - 1) / kSegBytes;
    if (p->zlen > 0x7FFFFFFFLL) return AM_ERR_ARG;
    PngArgs a;
    a.W = p->W; a.H = p->H; a.WPR = p->WPR; a.row_bytes = p->row_bytes; a.seg = p->seg; a.depth = depth;
    a.raw = p->raw; a.n_blocks = p->n_blocks; a.zlen = p->zlen; a.total = p->total;
    memcpy(a.head, p->head, 41);
    cudaStream_t st = (cudaStream_t)stream;
    // scratch: per frame n_seg x (32 lane offsets + segment bytes) unsigned, n_seg x 2 Adler sums, stream length, Adler-32
    static std::mutex mu;
    static std::map<int, std::pair<char*, size_t>> scratch;
    const size_t per_frame = (size_t)n_seg * (33 * sizeof(unsigned) + 2 * sizeof(unsigned long long)) + 16;
    const size_t need = (size_t)batch * per_frame + 64;
    char* base = nullptr;
    {
        int dev = 0;
        AM_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(mu);
        auto& sc = scratch[dev];
        if (sc.second < need) {
            if (sc.first) cudaFree(sc.first);
            sc.first = nullptr; sc.second = 0;
            AM_CUDA(cudaMalloc(&sc.first, need));
            sc.second = need;
        }
        base = sc.first;
    }
    unsigned long long* seg_adler = (unsigned long long*)base;
    long long* d_zlen = (long long*)(seg_adler + (size_t)batch * n_seg * 2);
    unsigned* lane_off = (unsigned*)(d_zlen + batch);
    unsigned* seg_bytes = lane_off + (size_t)batch * n_seg * 32;
    uint32_t* adler = seg_bytes + (size_t)batch * n_seg;
    AM_CUDA(cudaMemsetAsync(d_out, 0, (size_t)batch * p->total, st));       // the bit writer ORs into zeroed words
    const dim3 grid((unsigned)((n_seg + kDefWarps - 1) / kDefWarps), (unsigned)batch);
    k_png1_deflate_count<<<grid, kDefWarps * 32, 0, st>>>(d_bits, a, (int)n_seg, lane_off, seg_bytes, seg_adler);
    k_png1_deflate_scan<<<batch, 32, 0, st>>>(a, (int)n_seg, seg_bytes, seg_adler, adler, d_zlen, d_out);
    k_png1_deflate_emit<<<grid, kDefWarps * 32, 0, st>>>(d_bits, a, (int)n_seg, lane_off, seg_bytes, adler, d_out);
    k_png1_checksums<<<batch, kThreads, 0, st>>>(a, p->d_tables, nullptr, 0, d_out, d_zlen, d_sizes);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}

extern "C" int am_png1_encode_deflate(const uint32_t* d_bits, int batch, int height, int width, uint8_t* d_out, long long* d_sizes,
                                      void* stream) {
    return png_encode_deflate(d_bits, 1, batch, height, width, d_out, d_sizes, stream);
}
// 8-bit grayscale frames [batch][height][width] (e.g. the clean frames of stage 03 where two groups overlap and the pixel value is 254)
extern "C" long long am_png8_capacity(int width, int height) {
    if (width <= 0 || height <= 0) return 0;
    const long long raw = (long long)height * (1 + width);
    return ((8 + 25 + 12 + deflate_capacity(raw) + 12) + 15) & ~15LL;
}
extern "C" int am_png8_encode_deflate(const uint8_t* d_frames, int batch, int height, int width, uint8_t* d_out, long long* d_sizes,
                                      void* stream) {
    return png_encode_deflate((const uint32_t*)d_frames, 8, batch, height, width, d_out, d_sizes, stream);
}

extern "C" int am_png1_scanlines_to_bits(const uint8_t* d_scan, int batch, int height, int width, uint32_t* d_bits, void* stream) {
    if (!d_scan || !d_bits || batch <= 0 || width <= 0 || height <= 0) return AM_ERR_ARG;
    const int WPR = am_words_per_row_impl(width), rb = (width + 7) / 8;
    k_png1_scan_to_bits<<<dim3(am_div_up(WPR, 64), height, batch), 64, 0, (cudaStream_t)stream>>>(d_scan, height, width, WPR, rb, d_bits);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}
