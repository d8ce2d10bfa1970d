// cc_kernels.cu -- connected-component labeling, per-CC statistics, crops and temporal matching
// on bit-packed masks, hand-written for sm_100a (B200).  HBM/latency-bound integer work: coalesced
// word loads, warp-aggregated atomics, no tensor cores.
//
// Replaces (R/ = reference ACCESS2021_release/):
//   scipy.ndimage.label                      R/AccessMath/preprocessing/content/labeler.py:126
//   CC_AgeBoundaries                         R/accessmath_lib.c:357-413
//   Labeler.extractSpatioTemporalContent     R/AccessMath/preprocessing/content/labeler.py:117-191
//   IntervalIndex.find_matches + getOverlapFMeasure + CCStabilityEstimator.add_frame
//        R/AccessMath/preprocessing/tools/interval_index.py:42-99,
//        R/AM_CommonTools/data/connected_component.py:202-250,
//        R/AccessMath/preprocessing/content/cc_stability_estimator.py:41-155
//
// Data layout (all device, per frame f of a batch):
//   bits   [H][WPR] uint32, bit b of word w = pixel x = 32*w + b (LSB first), WPR = words/row padded to 4
//   runs   = maximal horizontal runs of ink, numbered in raster order (row, then x): run id order ==
//            raster order of first pixels, so the minimum run id of a component is its first raster pixel
//            and rank(root run) + 1 == the SciPy label.
//   labels [H][W] int32 (optional output)
//   label table (by label-1): min_y, max_y, min_x, max_x, count     (CC_AgeBoundaries outputs)
//   kept table  (CCs with count >= min_pixels, ascending label): raw label, crop offset
//   crops  : per kept CC a word-aligned bit-packed crop: rows min_y..max_y, words (min_x>>5)..(max_x>>5),
//            bits at their ABSOLUTE x position, so two crops AND together without shifting.
#include "am_common.cuh"
#include "../../include/accessmath_b200.h"
#include <cooperative_groups.h>
#include <cstdlib>
#include <cstring>

#define NONE_U32 0xFFFFFFFFu

struct CcFrame {                 // device pointers of one frame slot
    uint32_t* wfirst;            // [H][WPR] run slot of the first run that STARTS in word w (runs before it in its strip + strip base)
    uint32_t* run_comp;          // [NS*CAP] run slot -> strip-component slot
    int *c_min_x, *c_max_x, *c_min_y, *c_max_y, *c_count;   // [NS*CAP] strip-component partial statistics
    int* c_link;                 // [NS*CAP] union-find over strip components (seam merge), flattened by k_resolve
    int* c_label;                // [NS*CAP] final raster-order label of each strip component
    int2* strip_n;               // [NS] (runs, components) of each strip
    int* t_min_y; int* t_max_y; int* t_min_x; int* t_max_x; int* t_count;   // [ML]
    uint32_t* lab_crop_off;      // [ML]
    int* kept_label;             // [MK]
    uint32_t* kept_crop_off;     // [MK]
    uint32_t* crops;             // [CW]
    int* match_unique;           // [MK] result of temporal matching: unique idx per kept CC
};

struct am_cc_ctx {
    int W, H, WPR, B, ML, MK, CW, min_pixels;
    int R, NS, CAP;              // strip height (rows), strips per frame, run / component slots per strip (= R * ceil(W/2): worst case)
    int strip_smem;              // dynamic shared memory of k_strip_label
    CcFrame* h_frames;           // host copy of slot descriptors
    CcFrame* d_frames;           // device copy
    int* d_counts;               // [B][4]: n_runs, n_labels, n_kept, crop_words
    int* d_status;               // capacity-overflow flags
    void* slab;                  // one allocation
    int* h_counts;               // pinned
};

// ------------------------------------------------------------------------------------------------
// K0: uint8 mask -> bit-packed (ink = nonzero)
__global__ void k_pack_u8(const uint8_t* __restrict__ src, int W, int H, int WPR, uint32_t* __restrict__ bits) {
    const int f = blockIdx.z, y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint8_t* row = src + ((size_t)f * H + y) * W;
    bool ink = (x < W) && row[x] != 0;
    unsigned m = __ballot_sync(0xffffffffu, ink);
    if ((threadIdx.x & 31) == 0 && (x >> 5) < WPR) bits[((size_t)f * H + y) * WPR + (x >> 5)] = m;
}

// the same, and flags[f] |= 1 when frame f holds a value other than 0 / 255 (such a frame has no exact 1-bit form: the clean frames
// of stage 03 wrap to 254 where two groups overlap, cc_stability_estimator.py:660-661)
__global__ void k_pack_u8_exact(const uint8_t* __restrict__ src, int W, int H, int WPR, uint32_t* __restrict__ bits, int* __restrict__ flags) {
    const int f = blockIdx.z, y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint8_t* row = src + ((size_t)f * H + y) * W;
    const uint8_t v = x < W ? row[x] : 0;
    const unsigned m = __ballot_sync(0xffffffffu, v != 0);
    const unsigned odd = __ballot_sync(0xffffffffu, v != 0 && v != 255);
    if ((threadIdx.x & 31) == 0) {
        if ((x >> 5) < WPR) bits[((size_t)f * H + y) * WPR + (x >> 5)] = m;
        if (odd) atomicOr(&flags[f], 1);
    }
}

// bit-packed -> uint8 0/255 (used to hand masks back in the reference's format)
__global__ void k_unpack_u8(const uint32_t* __restrict__ bits, int W, int H, int WPR, uint8_t* __restrict__ dst) {
    const int f = blockIdx.z, y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    uint32_t m = bits[((size_t)f * H + y) * WPR + (x >> 5)];
    dst[((size_t)f * H + y) * W + x] = ((m >> (x & 31)) & 1u) ? 255 : 0;
}

__device__ __forceinline__ uint32_t mask_le(int b) { return (2u << b) - 1u; }   // bits 0..b

// union-find with "link the larger root under the smaller" (works on shared or global memory): the root of a set is
// its minimum id, i.e. the run / strip component that holds the first raster pixel
__device__ __forceinline__ int uf_find(int* parent, int a) {
    int p;
    while ((p = ((volatile int*)parent)[a]) != a) a = p;
    return a;
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a); b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }        // a > b : link the larger root under the smaller
        int old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;
    }
}

// One maximal run of ones of word m starting at the lowest set bit of x (x = the not yet visited bits of m)
__device__ __forceinline__ uint32_t next_segment(uint32_t m, uint32_t x, int* b0, int* len) {
    const int b = __ffs(x) - 1;
    const uint32_t t = ~(m >> b);                        // zero bits where the run continues
    const int l = t ? (__ffs(t) - 1) : 32;
    *b0 = b; *len = l;
    return (l == 32) ? 0xFFFFFFFFu : (((1u << l) - 1u) << b);
}

// ------------------------------------------------------------------------------------------------
// K1: one CTA per strip of R rows.  Everything a strip can decide alone happens in shared memory:
//   runs (maximal horizontal ink runs, numbered in raster order inside the strip) -> union-find over vertically
//   adjacent runs -> strip components (numbered by their first raster pixel) with bbox / pixel count reduced by
//   shared-memory atomics.  Global output: wfirst (one word per mask word), run_comp (one word per run) and one
//   row per strip component.  Slots are laid out strip by strip with a worst-case capacity, so global slot order ==
//   raster order of first pixels without any inter-CTA prefix.
#ifndef STRIP_THREADS
#define STRIP_THREADS 512
#endif
#ifndef STRIP_CAPS
#define STRIP_CAPS 2048          // strip components reduced in shared memory; the (rare) rest uses global atomics
#endif
#ifndef STRIP_MAX_RUNS
#define STRIP_MAX_RUNS 16384     // worst-case runs of a strip (R * ceil(W/2)) the shared-memory union-find is sized for
#endif
#ifndef STRIP_MIN_BLOCKS
#define STRIP_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(STRIP_THREADS, STRIP_MIN_BLOCKS)
k_strip_label(const uint32_t* __restrict__ bits, const CcFrame* __restrict__ frames, int W, int H, int WPR, int R, int CAP) {
    extern __shared__ uint32_t smem[];
    __shared__ int s_scan[33];
    __shared__ int s_rowbase[33];
    const int s = blockIdx.x, f = blockIdx.y;
    const int y0 = s * R, rows = min(R, H - y0), nwords = rows * WPR;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = STRIP_THREADS >> 5;
    uint32_t* s_bits = smem;
    int* s_first = (int*)(smem + R * WPR);
    int* s_parent = s_first + R * WPR;
    int* s_minx = s_parent + CAP;
    int* s_maxx = s_minx + STRIP_CAPS;
    int* s_miny = s_maxx + STRIP_CAPS;
    int* s_maxy = s_miny + STRIP_CAPS;
    int* s_cnt = s_maxy + STRIP_CAPS;
    const CcFrame fr = frames[f];
    const int base = s * CAP;                            // first run slot == first component slot of this strip

    // 1. the strip's mask words (rows are contiguous in memory)
    {
        const uint32_t* src = bits + ((size_t)f * H + y0) * WPR;
        if ((WPR & 3) == 0) {
            const uint4* s4 = (const uint4*)src; uint4* d4 = (uint4*)s_bits;
            for (int i = tid; i < (nwords >> 2); i += STRIP_THREADS) d4[i] = s4[i];
        } else {
            for (int i = tid; i < nwords; i += STRIP_THREADS) s_bits[i] = src[i];
        }
    }
    __syncthreads();
    // 2. warp per row: run starts per word, prefix inside the row
    for (int r = warp; r < rows; r += nwarps) {
        int running = 0;
        uint32_t carry = 0;
        for (int w0 = 0; w0 < WPR; w0 += 32) {
            const int w = w0 + lane;
            const uint32_t m = (w < WPR) ? s_bits[r * WPR + w] : 0u;
            uint32_t prev = __shfl_up_sync(0xffffffffu, m, 1);
            if (lane == 0) prev = carry;
            const int c = __popc(m & ~((m << 1) | (prev >> 31)));
            const int inc = warp_incl_scan(c);
            if (w < WPR) s_first[r * WPR + w] = running + inc - c;
            running += __shfl_sync(0xffffffffu, inc, 31);
            carry = __shfl_sync(0xffffffffu, m, 31);
        }
        if (lane == 0) s_rowbase[r + 1] = running;
    }
    __syncthreads();
    // 3. scan over the rows (R <= 32)
    if (warp == 0) {
        const int v = (lane < rows) ? s_rowbase[lane + 1] : 0;
        const int inc = warp_incl_scan(v);
        if (lane < rows) s_rowbase[lane + 1] = inc;
        if (lane == 0) s_rowbase[0] = 0;
    }
    __syncthreads();
    const int n_runs = s_rowbase[rows];
    // 4. strip-local run id of the first run starting in each word; parent = self
    for (int i = tid; i < nwords; i += STRIP_THREADS) {
        const int v = s_first[i] + s_rowbase[i / WPR];
        s_first[i] = v;
        fr.wfirst[(size_t)y0 * WPR + i] = (uint32_t)(base + v);
    }
    for (int i = tid; i < n_runs; i += STRIP_THREADS) s_parent[i] = i;
    __syncthreads();
    // 5. union the runs of vertically adjacent ink (one overlap-run start = one union)
    for (int i = tid + WPR; i < nwords; i += STRIP_THREADS) {
        const int w = i % WPR;
        const uint32_t m = s_bits[i], u = s_bits[i - WPR];
        const uint32_t ov = m & u;
        if (ov == 0) continue;
        const uint32_t mp = (w > 0) ? s_bits[i - 1] : 0u, upv = (w > 0) ? s_bits[i - WPR - 1] : 0u;
        uint32_t ovs = ov & ~(ov << 1);
        ovs &= ~(((mp & upv) >> 31) & 1u);              // overlap continuing from the previous word: done there
        if (ovs == 0) continue;
        const uint32_t st_m = m & ~((m << 1) | (mp >> 31));
        const uint32_t st_u = u & ~((u << 1) | (upv >> 31));
        const int fm = s_first[i], fu = s_first[i - WPR];
        while (ovs) {
            const int b = __ffs(ovs) - 1; ovs &= ovs - 1;
            // link the lower run straight to the upper run (no root search: chains are at most one hop per row);
            // a run that already had a different parent is a join of two upper runs -> a real union of those
            const int ia = fm + __popc(st_m & mask_le(b)) - 1, ib = fu + __popc(st_u & mask_le(b)) - 1;
            const int old = atomicMin(&s_parent[ia], ib);
            if (old != ia && old != ib) uf_union(s_parent, old, ib);
        }
    }
    __syncthreads();
    // 6. flatten: pointer jumping covers the one-hop-per-row chains (depth < R <= 32), then a plain root search
    //    for what joins left over
#pragma unroll 1
    for (int round = 0; round < 5; ++round) {
        for (int i = tid; i < n_runs; i += STRIP_THREADS) {
            const int p = s_parent[i];
            const int g = s_parent[p];
            if (g != p) s_parent[i] = g;
        }
        __syncthreads();
    }
    for (int i = tid; i < n_runs; i += STRIP_THREADS) {
        const int r = uf_find(s_parent, i);
        if (r != s_parent[i]) s_parent[i] = r;
    }
    __syncthreads();
    // 7. number the roots in run order (= raster order of first pixels); a root's parent entry becomes ~component
    int n_comp = 0;
    for (int i0 = 0; i0 < n_runs; i0 += STRIP_THREADS) {
        const int i = i0 + tid;
        const int isroot = (i < n_runs) && (s_parent[i] == i);
        int tot;
        const int cid = n_comp + block_excl_scan(isroot, s_scan, &tot);
        if (isroot) {
            s_parent[i] = ~cid;
            if (cid < STRIP_CAPS) { s_minx[cid] = W; s_maxx[cid] = 0; s_miny[cid] = H; s_maxy[cid] = 0; s_cnt[cid] = 0; }
            else {
                const int slot = base + cid;
                fr.c_min_x[slot] = W; fr.c_max_x[slot] = 0; fr.c_min_y[slot] = H; fr.c_max_y[slot] = 0; fr.c_count[slot] = 0;
                fr.c_link[slot] = slot;
            }
        }
        n_comp += tot;
    }
    __syncthreads();
    // 8. every segment (part of a run inside one word) adds to its component; run starts record run -> component
    for (int i = tid; i < nwords; i += STRIP_THREADS) {
        const uint32_t m = s_bits[i];
        if (m == 0) continue;
        const int r = i / WPR, w = i - r * WPR, y = y0 + r;
        const uint32_t cin = (w > 0) ? (s_bits[i - 1] >> 31) : 0u;
        const uint32_t starts = m & ~((m << 1) | cin);
        const int first = s_first[i];
        uint32_t x = m;
        while (x) {
            int b0, len;
            const uint32_t seg = next_segment(m, x, &b0, &len);
            x &= ~seg;
            const int id = first + __popc(starts & mask_le(b0)) - 1;
            const int p = s_parent[id];
            const int cid = (p < 0) ? ~p : ~s_parent[p];
            const int xs = w * 32 + b0, xe = xs + len - 1;
            if (cid < STRIP_CAPS) {
                atomicMin(&s_minx[cid], xs); atomicMax(&s_maxx[cid], xe); atomicMax(&s_maxy[cid], y); atomicAdd(&s_cnt[cid], len);
                if (p < 0) s_miny[cid] = y;              // the root run holds the first raster pixel: its row is min_y
            } else {
                const int slot = base + cid;
                atomicMin(&fr.c_min_x[slot], xs); atomicMax(&fr.c_max_x[slot], xe);
                atomicMin(&fr.c_min_y[slot], y); atomicMax(&fr.c_max_y[slot], y); atomicAdd(&fr.c_count[slot], len);
            }
            if (seg & starts) fr.run_comp[base + id] = (uint32_t)(base + cid);
        }
    }
    __syncthreads();
    // 9. strip-component rows
    for (int c = tid; c < min(n_comp, STRIP_CAPS); c += STRIP_THREADS) {
        const int slot = base + c;
        fr.c_min_x[slot] = s_minx[c]; fr.c_max_x[slot] = s_maxx[c]; fr.c_min_y[slot] = s_miny[c]; fr.c_max_y[slot] = s_maxy[c];
        fr.c_count[slot] = s_cnt[c]; fr.c_link[slot] = slot;
    }
    if (tid == 0) fr.strip_n[s] = make_int2(n_runs, n_comp);
}

// K2: one CTA per frame over its strip components (a few thousand): union the components of vertically adjacent ink
// across the seams between strips, flatten, number the roots in slot order (= raster order of first pixels = the
// SciPy label), reduce the partial statistics into the label table (= CC_AgeBoundaries, accessmath_lib.c:357-413),
// then compact the labels with count >= min_pixels (labeler.py:177), lay out their crops and zero the crop arena.
#define RESOLVE_THREADS 1024
__global__ void __launch_bounds__(RESOLVE_THREADS)
k_resolve(const uint32_t* __restrict__ bits, const CcFrame* __restrict__ frames, int* __restrict__ counts, int H, int WPR, int R,
          int NS, int CAP, int ML, int MK, int CW, int min_pixels, int* __restrict__ status) {
    extern __shared__ int s_prefix[];                    // [NS + 1] components before each strip
    __shared__ int sm[33];
    const int f = blockIdx.x, tid = threadIdx.x;
    const CcFrame fr = frames[f];
    // 0. seams: one (seam, word) per thread
    for (int i = tid; i < (NS - 1) * WPR; i += RESOLVE_THREADS) {
        const int sidx = i / WPR, w = i - sidx * WPR, y = (sidx + 1) * R;
        const uint32_t* row = bits + ((size_t)f * H + y) * WPR;
        const uint32_t* up = row - WPR;
        const uint32_t m = row[w], u = up[w];
        const uint32_t ov = m & u;
        if (ov == 0) continue;
        const uint32_t mp = (w > 0) ? row[w - 1] : 0u, upv = (w > 0) ? up[w - 1] : 0u;
        uint32_t ovs = ov & ~(ov << 1);
        ovs &= ~(((mp & upv) >> 31) & 1u);              // overlap continuing from the previous word: done there
        if (ovs == 0) continue;
        const uint32_t st_m = m & ~((m << 1) | (mp >> 31));
        const uint32_t st_u = u & ~((u << 1) | (upv >> 31));
        const int fm = (int)fr.wfirst[(size_t)y * WPR + w], fu = (int)fr.wfirst[(size_t)(y - 1) * WPR + w];
        while (ovs) {
            const int b = __ffs(ovs) - 1; ovs &= ovs - 1;
            const int ca = (int)fr.run_comp[fm + __popc(st_m & mask_le(b)) - 1];
            const int cb = (int)fr.run_comp[fu + __popc(st_u & mask_le(b)) - 1];
            uf_union(fr.c_link, ca, cb);
        }
    }
    int total = 0, runs = 0;
    for (int s0 = 0; s0 < NS; s0 += RESOLVE_THREADS) {
        const int s = s0 + tid;
        const int2 n = (s < NS) ? fr.strip_n[s] : make_int2(0, 0);
        int tot, rtot;
        const int ex = block_excl_scan(n.y, sm, &tot);
        block_excl_scan(n.x, sm, &rtot);
        if (s < NS) s_prefix[s] = total + ex;
        total += tot; runs += rtot;
    }
    if (tid == 0) s_prefix[NS] = total;
    __syncthreads();
    auto slot_of = [&](int idx) {                        // compact component index -> slot
        int lo = 0, hi = NS;                             // last strip with prefix <= idx
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_prefix[mid] <= idx) lo = mid; else hi = mid; }
        return lo * CAP + (idx - s_prefix[lo]);
    };
    // a. flatten
    for (int idx = tid; idx < total; idx += RESOLVE_THREADS) {
        const int c = slot_of(idx);
        fr.c_link[c] = uf_find(fr.c_link, c);
    }
    __syncthreads();
    // b. roots -> labels, own statistics start the table row
    int n_labels = 0;
    for (int i0 = 0; i0 < total; i0 += RESOLVE_THREADS) {
        const int idx = i0 + tid;
        int c = 0, isroot = 0;
        if (idx < total) { c = slot_of(idx); isroot = (fr.c_link[c] == c); }
        int tot;
        const int lab = n_labels + block_excl_scan(isroot, sm, &tot);       // 0-based
        if (isroot) {
            fr.c_label[c] = lab + 1;
            if (lab < ML) {
                fr.t_min_y[lab] = fr.c_min_y[c]; fr.t_max_y[lab] = fr.c_max_y[c]; fr.t_min_x[lab] = fr.c_min_x[c];
                fr.t_max_x[lab] = fr.c_max_x[c]; fr.t_count[lab] = fr.c_count[c];
            }
        }
        n_labels += tot;
    }
    __syncthreads();
    // c. the other strip components of a label (components that cross a seam) fold into the root's row
    for (int idx = tid; idx < total; idx += RESOLVE_THREADS) {
        const int c = slot_of(idx);
        const int r = fr.c_link[c];
        if (r == c) continue;
        const int lab = fr.c_label[r];
        fr.c_label[c] = lab;
        if (lab <= ML) {
            const int l = lab - 1;                        // min_y: the root holds the first raster pixel
            atomicMax(&fr.t_max_y[l], fr.c_max_y[c]); atomicMin(&fr.t_min_x[l], fr.c_min_x[c]);
            atomicMax(&fr.t_max_x[l], fr.c_max_x[c]); atomicAdd(&fr.t_count[l], fr.c_count[c]);
        }
    }
    __syncthreads();
    // d. kept labels + crop offsets
    const int n = min(n_labels, ML);
    int kcarry = 0; unsigned wcarry = 0; bool over = false;
    for (int i0 = 0; i0 < n; i0 += RESOLVE_THREADS) {
        const int l = i0 + tid;
        int keep = 0, words = 0;
        if (l < n && fr.t_count[l] >= min_pixels) {
            keep = 1;
            words = ((fr.t_max_x[l] >> 5) - (fr.t_min_x[l] >> 5) + 1) * (fr.t_max_y[l] - fr.t_min_y[l] + 1);
        }
        int ktot, wtot;
        const int kex = block_excl_scan(keep, sm, &ktot);
        const int wex = block_excl_scan(words, sm, &wtot);
        if (l < n) {
            uint32_t off = NONE_U32;
            if (keep) {
                const int ki = kcarry + kex;
                const unsigned wo = wcarry + (unsigned)wex;
                if (ki < MK && wo + (unsigned)words <= (unsigned)CW) {
                    fr.kept_label[ki] = l + 1; fr.kept_crop_off[ki] = wo; fr.match_unique[ki] = -1; off = wo;
                } else over = true;
            }
            fr.lab_crop_off[l] = off;
        }
        kcarry += ktot; wcarry += (unsigned)wtot;
    }
    const int any_over = __syncthreads_or(over);
    {                                                    // zero the crop arena (16-byte aligned, padded by 4 words)
        const int n4 = (int)((min(wcarry, (unsigned)CW) + 3u) >> 2);
        uint4* c4 = (uint4*)fr.crops;
        for (int i = tid; i < n4; i += RESOLVE_THREADS) c4[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        if (any_over) atomicOr(status, 4);
        if (n_labels > ML) atomicOr(status, 2);
        counts[f * 4 + 0] = runs; counts[f * 4 + 1] = n_labels;
        counts[f * 4 + 2] = min(kcarry, MK); counts[f * 4 + 3] = (int)min(wcarry, (unsigned)CW);
    }
}

__device__ __forceinline__ int label_of_run(const CcFrame& fr, int run_slot) { return fr.c_label[fr.run_comp[run_slot]]; }

// K3: one thread per mask word: OR every segment into its CC's crop (labeler.py:183, bit-packed, absolute x)
__global__ void k_crop_fill(const uint32_t* __restrict__ bits, const CcFrame* __restrict__ frames, int H, int WPR) {
    const int f = blockIdx.y;
    const uint32_t* fbits = bits + (size_t)f * H * WPR;
    const int nwords = H * WPR;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += gridDim.x * blockDim.x) {
        const uint32_t m = fbits[i];
        if (m == 0) continue;
        const int y = i / WPR, w = i - y * WPR;
        const CcFrame fr = frames[f];
        const uint32_t cin = (w > 0) ? (fbits[i - 1] >> 31) : 0u;
        const uint32_t starts = m & ~((m << 1) | cin);
        const int first = (int)fr.wfirst[i];
        uint32_t x = m;
        while (x) {
            int b0, len;
            const uint32_t seg = next_segment(m, x, &b0, &len);
            x &= ~seg;
            const int l = label_of_run(fr, first + __popc(starts & mask_le(b0)) - 1) - 1;
            const uint32_t off = fr.lab_crop_off[l];
            if (off == NONE_U32) continue;
            const int wx0 = fr.t_min_x[l] >> 5, cw = (fr.t_max_x[l] >> 5) - wx0 + 1;
            atomicOr(&fr.crops[off + (size_t)(y - fr.t_min_y[l]) * cw + (w - wx0)], seg);
        }
    }
}

// K4 (optional): label image (int32, 0 = background = scipy.ndimage.label's output), four pixels per thread
__global__ void k_label_image(const uint32_t* __restrict__ bits, const CcFrame* __restrict__ frames, int W, int H, int WPR,
                              int32_t* __restrict__ labels) {
    const int f = blockIdx.y;
    const uint32_t* fbits = bits + (size_t)f * H * WPR;
    int32_t* out = labels + (size_t)f * H * W;
    const int gpr = (W + 3) >> 2, ngroups = H * gpr;     // groups of 4 pixels per row
    const bool vec = (W & 3) == 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ngroups; i += gridDim.x * blockDim.x) {
        const int y = i / gpr, x0 = (i - y * gpr) << 2, w = x0 >> 5, b = x0 & 31;
        const uint32_t m = fbits[(size_t)y * WPR + w];
        int lab[4] = {0, 0, 0, 0};
        if ((m >> b) & 15u) {
            const CcFrame fr = frames[f];
            const uint32_t cin = (w > 0) ? (fbits[(size_t)y * WPR + w - 1] >> 31) : 0u;
            const uint32_t starts = m & ~((m << 1) | cin);
            const int first = (int)fr.wfirst[(size_t)y * WPR + w];
            int prev_id = -1, prev_lab = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if ((m >> (b + k)) & 1u) {
                    const int id = first + __popc(starts & mask_le(b + k)) - 1;
                    if (id != prev_id) { prev_id = id; prev_lab = label_of_run(fr, id); }
                    lab[k] = prev_lab;
                }
            }
        }
        int32_t* dst = out + (size_t)y * W + x0;
        if (vec) *(int4*)dst = make_int4(lab[0], lab[1], lab[2], lab[3]);
        else {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (x0 + k < W) dst[k] = lab[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Legacy operator: CC_AgeBoundaries on an arbitrary int32 label image (+ fp32 ages), R/accessmath_lib.c:357-413
__device__ __forceinline__ unsigned f32_key(float v) {     // order-preserving float -> unsigned
    unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_f32(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// CC_AgeBoundaries' age rule (accessmath_lib.c:405-407) is `if (age < 0 || a < age) age = a` over the pixels of a component in
// raster order.  With L = the raster index of the component's LAST pixel whose age is negative (none: no L) that sequence
// ends as  min(a_i : i > L)  or, when L is the component's last pixel, a_L itself -- so two order-free reductions reproduce
// it exactly: pass 1 (k_ageb_scan) reduces bounds / count / plain minimum and atomicMax-es L, pass 2 (k_ageb_neg, a no-op
// unless some age was negative) redoes the minimum over the pixels behind L.
__global__ void k_ageb_init(int n, int W, int H, int* mny, int* mxy, int* mnx, int* mxx, int* cnt, unsigned* agek, long long* lastneg,
                            int* any_neg) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *any_neg = 0;
    if (i >= n) return;
    mny[i] = H; mxy[i] = 0; mnx[i] = W; mxx[i] = 0; cnt[i] = 0; agek[i] = 0xFFFFFFFFu; lastneg[i] = 0;
}
__global__ void k_ageb_scan(const int32_t* __restrict__ labels, const float* __restrict__ ages, int W, int H, int n,
                            int* mny, int* mxy, int* mnx, int* mxx, int* cnt, unsigned* agek, long long* lastneg, int* any_neg) {
    const int y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    int lab = 0; unsigned ak = 0xFFFFFFFFu; long long neg = -1;
    if (x < W) {
        size_t idx = (size_t)y * W + x;
        lab = labels[idx];
        if (lab < 0 || lab > n) lab = 0;                 // the reference would write out of bounds here
        if (lab > 0 && ages) { const float a = ages[idx]; ak = f32_key(a); if (a < 0.0f) neg = (long long)idx; }
        else if (lab > 0) ak = f32_key(0.0f);
    }
    unsigned peers = __match_any_sync(0xffffffffu, lab);
    int mn_x = __reduce_min_sync(peers, x), mx_x = __reduce_max_sync(peers, x);
    int c = __popc(peers);
    unsigned amin = __reduce_min_sync(peers, ak);
    int negx = __reduce_max_sync(peers, neg >= 0 ? x : -1);
    if (lab > 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) {
        int l = lab - 1;
        atomicMin(&mnx[l], mn_x); atomicMax(&mxx[l], mx_x);
        atomicMin(&mny[l], y); atomicMax(&mxy[l], y);
        atomicAdd(&cnt[l], c);
        atomicMin(&agek[l], amin);
        if (negx >= 0) { atomicMax((unsigned long long*)&lastneg[l], (unsigned long long)((long long)y * W + negx + 1)); *any_neg = 1; }
    }
}
// lastneg holds (raster index + 1) of the last negative-age pixel through the unsigned atomicMax above (0 / -1 = none)
__global__ void k_ageb_neg_reset(int n, const long long* lastneg, unsigned* agek, const int* any_neg) {
    if (!*any_neg) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && lastneg[i] > 0) agek[i] = 0xFFFFFFFFu;
}
__global__ void k_ageb_neg(const int32_t* __restrict__ labels, const float* __restrict__ ages, int W, int H, int n,
                           const long long* lastneg, unsigned* agek, const int* any_neg) {
    if (!*any_neg) return;
    const int y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const long long idx = (long long)y * W + x;
    const int lab = labels[idx];
    if (lab <= 0 || lab > n) return;
    const long long L = lastneg[lab - 1];
    if (L > 0 && idx + 1 > L) atomicMin(&agek[lab - 1], f32_key(ages[idx]));
}
__global__ void k_ageb_finish(int n, const int* cnt, const unsigned* agek, const long long* lastneg, const float* __restrict__ ages,
                              float* out_age) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = -1.0f;                                          // accessmath_lib.c:373 "-1 = unset"
    if (cnt[i] > 0) a = (agek[i] == 0xFFFFFFFFu && lastneg[i] > 0) ? ages[lastneg[i] - 1] : key_f32(agek[i]);
    out_age[i] = a;
}
// labels / ages / outputs in DEVICE memory; d_tab = scratch of 10 * n ints.  Asynchronous on `st`.
static int ageb_run(const int32_t* d_lab, const float* d_age, int width, int height, int n, int* d_tab, cudaStream_t st) {
    int *mny = d_tab, *mxy = d_tab + n, *mnx = d_tab + 2 * n, *mxx = d_tab + 3 * n, *cnt = d_tab + 4 * n;
    unsigned* agek = (unsigned*)(d_tab + 5 * n); float* oage = (float*)(d_tab + 6 * n);
    long long* lastneg = (long long*)(d_tab + 7 * n + (n & 1));              // 8-byte aligned (d_tab is)
    int* any_neg = d_tab + 10 * n + 1;
    k_ageb_init<<<am_div_up(n, 256), 256, 0, st>>>(n, width, height, mny, mxy, mnx, mxx, cnt, agek, lastneg, any_neg);
    k_ageb_scan<<<dim3(am_div_up(width, 256), height), 256, 0, st>>>(d_lab, d_age, width, height, n, mny, mxy, mnx, mxx, cnt, agek, lastneg, any_neg);
    if (d_age) {
        k_ageb_neg_reset<<<am_div_up(n, 256), 256, 0, st>>>(n, lastneg, agek, any_neg);
        k_ageb_neg<<<dim3(am_div_up(width, 256), height), 256, 0, st>>>(d_lab, d_age, width, height, n, lastneg, agek, any_neg);
    }
    k_ageb_finish<<<am_div_up(n, 256), 256, 0, st>>>(n, cnt, agek, lastneg, d_age, oage);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}

// ------------------------------------------------------------------------------------------------
// Temporal matching (CCStabilityEstimator.add_frame, cc_stability_estimator.py:41-155)
struct am_estimator {
    int fused_mode = -1, fused_grid = 0;     // k_match_fused launch configuration of the device this estimator lives on
    int W, H, max_gap, MU, MA;
    unsigned long long AW;       // arena capacity in words
    unsigned long long* tmp_off; // [MA] scratch for export/import
    double min_recall, min_precision;
    // unique table (index = global unique idx)
    int *u_min_x, *u_max_x, *u_min_y, *u_max_y, *u_size, *u_last, *u_first_frame, *u_first_label;
    unsigned long long* u_crop_off;
    uint32_t* arena;             // first-seen crops, append only
    int *act[2];                 // active lists (unordered), double buffered
    int* newlist;                // [MA] current-CC indices of the frame's new uniques (k_match_update -> k_match_copy)
    uint2* box[2];               // packed bbox of act[i]: .x = min_x | min_y << 16, .y = (max_x | max_y << 16) | 0x80008000
    int cur;                     // which act / box buffer is current
    // device scalars: [0]=n_uniq [1]=n_act [2]=img_idx [3]=status [4]=n_pairs [5]=n_items [6]=block ticket
    //                 [7]=n_act being built [8]=new uniques of the frame ; 64-bit: tested, arena_used
    int* d_scal; unsigned long long* d_scal64;
    // per-frame candidate work lists (reset by k_match_update): scal[4] = n_pairs, scal[5] = n_items
    int MP, MI;                  // capacities
    int2* pair_cu;               // [MP] (current kept index, unique index) of every bbox-overlapping candidate
    int* pair_m;                 // [MP] pixel overlap of the pair (popcount of AND), accumulated by k_match_overlap
    int2* items;                 // [MI] (pair index, chunk of MATCH_CHUNK words) -- large overlaps are split over warps
    void* slab;
    int* h_scal; unsigned long long* h_scal64;   // pinned
};
#ifndef MATCH_CHUNK
#define MATCH_CHUNK 512          // words of one overlap work item (one warp): a board-sized pair (64k words) spreads over 128 warps
#endif
#define MATCH_NONE 0x7fffffff

// Temporal matching is pair-parallel.  The reference tests a CC's candidates in ascending unique index and stops
// computing at the first one that passes (cc_stability_estimator.py:90-108) while still COUNTING every candidate
// (tempo_count, :85).  Equivalent and parallel: compute the overlap of every candidate pair, then take the MINIMUM
// unique index among the passing ones (so the order of the active list is irrelevant).  Large overlaps (a board-sized
// CC) are split into MATCH_CHUNK-word items so that one giant pair does not serialise on a single warp.
//
// M1a: all-pairs inclusive bbox test (= the two IntervalIndex sweeps + set intersection, interval_index.py:42-99,
// cc_stability_estimator.py:73-84), tiled: a block tests MP_THREADS current CCs (one per thread) against a tile of
// MP_TILE active boxes staged in shared memory.  Boxes are packed 2 x 16 bit with a bias bit so that one subtraction
// compares x and y at once: ((C1 | bias) - lo) has bit 15 / 31 set  <=>  c.max_x >= u.min_x / c.max_y >= u.min_y.
#define MP_TILE 256
#define MP_THREADS 256
#define BOX_BIAS 0x80008000u
#define MP_QUEUE 2048            // candidate pairs a tile queues in shared memory (the rest appends directly)
__device__ __forceinline__ void append_pair(int c, int u, int cx0, int cx1, int cy0, int cy1, uint2 b, int* scal, int2* pair_cu,
                                            int* pair_m, int2* items, int MP, int MI) {
    const int ux0 = (int)(b.x & 0xFFFFu), uy0 = (int)(b.x >> 16), ux1 = (int)(b.y & 0x7FFFu), uy1 = (int)((b.y >> 16) & 0x7FFFu);
    const int nw = (min(cx1, ux1) >> 5) - (max(cx0, ux0) >> 5) + 1;
    const int tot = nw * (min(cy1, uy1) - max(cy0, uy0) + 1);
    const int n_items = (tot + MATCH_CHUNK - 1) / MATCH_CHUNK;
    const int pi = atomicAdd(&scal[4], 1);
    int ii = atomicAdd(&scal[5], n_items);
    if (pi < MP) {
        pair_cu[pi] = make_int2(c, u); pair_m[pi] = 0;
        for (int k = 0; k < n_items; ++k, ++ii)
            if (ii < MI) items[ii] = make_int2(pi, k);
    }
}
// (bodies take a virtual block index / block count so that the per-frame kernels and the fused cooperative kernel share them;
//  nothing that another phase of the same frame writes is read through const __restrict__ -- no ld.global.nc inside the fused kernel)
__device__ __forceinline__ void match_pairs_body(int bid, int nblk, const CcFrame* frames, const int* __restrict__ counts, int f, const int* act,
              const uint2* box, int* scal, int2* pair_cu, int* pair_m,
              int2* items, int MP, int MI, unsigned long long* tested_total) {
    __shared__ __align__(16) uint2 s_box[MP_TILE];
    __shared__ int4 s_cc[MP_THREADS];                    // boxes of the tile's current CCs
    __shared__ unsigned s_queue[MP_QUEUE];               // (thread << 16) | j
    __shared__ int s_nq, s_base[2], s_scan[33];
    const CcFrame fr = frames[f];
    const int n_kept = counts[f * 4 + 2];
    const int img_idx = scal[2];
    const int n_act = (img_idx == 0) ? 0 : scal[1];      // frame 0: every CC becomes a unique CC (:52-69)
    const int tiles_c = (n_kept + MP_THREADS - 1) / MP_THREADS, tiles_a = max(1, (n_act + MP_TILE - 1) / MP_TILE);
    const int tid = threadIdx.x;
    unsigned hits = 0;
    for (int t = bid; t < tiles_c * tiles_a; t += nblk) {
        const int ci = t % tiles_c, ai = t / tiles_c;
        const int a0 = ai * MP_TILE, jn = min(MP_TILE, n_act - a0);
        __syncthreads();
        if (tid == 0) s_nq = 0;
        for (int j = tid; j < MP_TILE; j += MP_THREADS)      // padding boxes (min = 32767, max = 0) never overlap
            s_box[j] = (j < jn) ? box[a0 + j] : make_uint2(0x7FFF7FFFu, BOX_BIAS);
        const int c = ci * MP_THREADS + tid;
        int cx0 = 0, cx1 = 0, cy0 = 0, cy1 = 0;
        if (c < n_kept) {
            if (ai == 0) fr.match_unique[c] = MATCH_NONE;
            const int l = fr.kept_label[c] - 1;
            cx0 = fr.t_min_x[l]; cx1 = fr.t_max_x[l]; cy0 = fr.t_min_y[l]; cy1 = fr.t_max_y[l];
        }
        s_cc[tid] = make_int4(cx0, cx1, cy0, cy1);
        __syncthreads();
        if (c < n_kept && jn > 0) {
            const uint32_t C0 = (uint32_t)cx0 | ((uint32_t)cy0 << 16), C1 = ((uint32_t)cx1 | ((uint32_t)cy1 << 16)) | BOX_BIAS;
            const int jn4 = (jn + 3) & ~3;
#pragma unroll 2
            for (int j = 0; j < jn4; j += 4) {               // four boxes per step, branch only on a hit
                const uint4 b01 = *(const uint4*)&s_box[j], b23 = *(const uint4*)&s_box[j + 2];
                const uint32_t t0 = (C1 - b01.x) & (b01.y - C0), t1 = (C1 - b01.z) & (b01.w - C0);
                const uint32_t t2 = (C1 - b23.x) & (b23.y - C0), t3 = (C1 - b23.z) & (b23.w - C0);
                unsigned m = ((t0 & BOX_BIAS) == BOX_BIAS ? 1u : 0u) | ((t1 & BOX_BIAS) == BOX_BIAS ? 2u : 0u) |
                             ((t2 & BOX_BIAS) == BOX_BIAS ? 4u : 0u) | ((t3 & BOX_BIAS) == BOX_BIAS ? 8u : 0u);
                while (m) {
                    const int jj = j + __ffs(m) - 1; m &= m - 1;
                    ++hits;
                    const int q = atomicAdd(&s_nq, 1);
                    if (q < MP_QUEUE) s_queue[q] = ((unsigned)tid << 16) | (unsigned)jj;
                    else append_pair(c, act[a0 + jj], cx0, cx1, cy0, cy1, s_box[jj], scal, pair_cu, pair_m, items, MP, MI);
                }
            }
        }
        __syncthreads();
        // queued candidates: one reservation of pair / item slots per round of MP_THREADS entries (the two list
        // counters are single addresses: per-candidate atomics would serialise in L2)
        const int nq = min(s_nq, MP_QUEUE);
        for (int q0 = 0; q0 < nq; q0 += MP_THREADS) {
            const int q = q0 + tid;
            int n_items = 0, cc_i = 0, u = 0;
            if (q < nq) {
                const unsigned e = s_queue[q];
                const int ct = (int)(e >> 16), j = (int)(e & 0xFFFFu);
                const int4 cc = s_cc[ct];
                const uint2 b = s_box[j];
                const int ux0 = (int)(b.x & 0xFFFFu), uy0 = (int)(b.x >> 16), ux1 = (int)(b.y & 0x7FFFu), uy1 = (int)((b.y >> 16) & 0x7FFFu);
                const int nw = (min(cc.y, ux1) >> 5) - (max(cc.x, ux0) >> 5) + 1;
                n_items = (nw * (min(cc.w, uy1) - max(cc.z, uy0) + 1) + MATCH_CHUNK - 1) / MATCH_CHUNK;
                cc_i = ci * MP_THREADS + ct; u = act[a0 + j];
            }
            int tot;
            const int ex = block_excl_scan(n_items, s_scan, &tot);
            if (tid == 0) { s_base[0] = atomicAdd(&scal[4], min(MP_THREADS, nq - q0)); s_base[1] = atomicAdd(&scal[5], tot); }
            __syncthreads();
            if (q < nq) {
                const int pi = s_base[0] + tid;
                int ii = s_base[1] + ex;
                if (pi < MP) {
                    pair_cu[pi] = make_int2(cc_i, u); pair_m[pi] = 0;
                    for (int k = 0; k < n_items; ++k, ++ii)
                        if (ii < MI) items[ii] = make_int2(pi, k);
                }
            }
        }
    }
    hits = __reduce_add_sync(0xffffffffu, hits);
    if ((tid & 31) == 0 && hits) atomicAdd(tested_total, (unsigned long long)hits);
}
__global__ void __launch_bounds__(MP_THREADS, 4)
k_match_pairs(const CcFrame* frames, const int* __restrict__ counts, int f, const int* act, const uint2* box, int* scal, int2* pair_cu,
              int* pair_m, int2* items, int MP, int MI, unsigned long long* tested_total) {
    match_pairs_body(blockIdx.x, gridDim.x, frames, counts, f, act, box, scal, pair_cu, pair_m, items, MP, MI, tested_total);
}

// M1b: persistent warps over the work items: popcount(cc.mask & unique.mask) on the bit-packed crops, which sit at
// their ABSOLUTE x position so the AND needs no shifting (connected_component.py:211-228)
__device__ __forceinline__ void match_overlap_body(int bid, int nblk, const CcFrame* frames, int f, int* scal,
                                const int* u_min_x, const int* u_max_x, const int* u_min_y, const int* u_max_y,
                                const unsigned long long* u_crop_off, const uint32_t* arena, const int2* pair_cu, int* pair_m,
                                const int2* items, int MP, int MI) {
    const CcFrame fr = frames[f];
    const int lane = threadIdx.x & 31;
    const int wid = (bid * blockDim.x + threadIdx.x) >> 5, nwarps = (nblk * blockDim.x) >> 5;
    const int n_items = min(scal[5], MI);
    for (int it = wid; it < n_items; it += nwarps) {
        const int2 item = items[it];
        if (item.x >= MP) continue;
        const int2 cu = pair_cu[item.x];
        const int l = fr.kept_label[cu.x] - 1;
        const int cx0 = fr.t_min_x[l], cx1 = fr.t_max_x[l], cy0 = fr.t_min_y[l], cy1 = fr.t_max_y[l];
        const int ux0 = u_min_x[cu.y], ux1 = u_max_x[cu.y], uy0 = u_min_y[cu.y], uy1 = u_max_y[cu.y];
        const int cwx0 = cx0 >> 5, ccw = (cx1 >> 5) - cwx0 + 1;
        const int uwx0 = ux0 >> 5, ucw = (ux1 >> 5) - uwx0 + 1;
        const uint32_t* ccrop = fr.crops + fr.kept_crop_off[cu.x];
        const uint32_t* ucrop = arena + u_crop_off[cu.y];
        const int y0 = max(cy0, uy0), y1 = min(cy1, uy1);
        const int w0 = max(cwx0, uwx0), w1 = min(cx1 >> 5, ux1 >> 5);
        const int nw = w1 - w0 + 1, tot = nw * (y1 - y0 + 1);
        const int i0 = item.y * MATCH_CHUNK, i1 = min(tot, i0 + MATCH_CHUNK);
        int m = 0;
#pragma unroll 4
        for (int i = i0 + lane; i < i1; i += 32) {
            const int ry = i / nw, ww = w0 + (i - ry * nw), yy = y0 + ry;
            const uint32_t a = ccrop[(size_t)(yy - cy0) * ccw + (ww - cwx0)];
            const uint32_t b = ucrop[(size_t)(yy - uy0) * ucw + (ww - uwx0)];
            m += __popc(a & b);
        }
        m = __reduce_add_sync(0xffffffffu, m);
        if (lane == 0 && m) atomicAdd(&pair_m[item.x], m);
    }
}
__global__ void k_match_overlap(const CcFrame* frames, int f, int* scal, const int* u_min_x, const int* u_max_x, const int* u_min_y,
                                const int* u_max_y, const unsigned long long* u_crop_off, const uint32_t* arena, const int2* pair_cu,
                                int* pair_m, const int2* items, int MP, int MI) {
    match_overlap_body(blockIdx.x, gridDim.x, frames, f, scal, u_min_x, u_max_x, u_min_y, u_max_y, u_crop_off, arena, pair_cu, pair_m, items, MP, MI);
}

// M1c: per pair: recall / precision in IEEE fp64 (connected_component.py:239-240), lowest passing unique index wins
__device__ __forceinline__ void match_select_body(int bid, int nblk, const CcFrame* frames, int f, const int* scal, const int* u_size,
                               const int2* pair_cu, const int* pair_m, int MP, double min_recall, double min_precision) {
    const CcFrame fr = frames[f];
    const int n_pairs = min(scal[4], MP);
    for (int i = bid * blockDim.x + threadIdx.x; i < n_pairs; i += nblk * blockDim.x) {
        const int2 cu = pair_cu[i];
        const int m = pair_m[i];
        const int csz = fr.t_count[fr.kept_label[cu.x] - 1];
        const double recall = (double)m / (double)csz;
        const double precision = (double)m / (double)u_size[cu.y];
        if (recall >= min_recall && precision >= min_precision) atomicMin(&fr.match_unique[cu.x], cu.y);
    }
}
__global__ void k_match_select(const CcFrame* frames, int f, const int* scal, const int* u_size, const int2* pair_cu, const int* pair_m, int MP,
                               double min_recall, double min_precision) {
    match_select_body(blockIdx.x, gridDim.x, frames, f, scal, u_size, pair_cu, pair_m, MP, min_recall, min_precision);
}

// M2a: matched CCs refresh their unique's last-seen frame (:104) -- must be complete before the expiry pass reads it
__device__ __forceinline__ void match_refresh_body(int bid, int nblk, const CcFrame* frames, const int* __restrict__ counts, int f, const int* scal,
                                int* u_last) {
    const CcFrame fr = frames[f];
    const int n_kept = counts[f * 4 + 2], img_idx = scal[2];
    for (int c = bid * blockDim.x + threadIdx.x; c < n_kept; c += nblk * blockDim.x) {
        const int u = fr.match_unique[c];
        if (u != MATCH_NONE) u_last[u] = img_idx;
    }
}
__global__ void k_match_refresh(const CcFrame* frames, const int* __restrict__ counts, int f, const int* scal, int* u_last) {
    match_refresh_body(blockIdx.x, gridDim.x, frames, counts, f, scal, u_last);
}

__device__ __forceinline__ uint2 pack_box(int x0, int x1, int y0, int y1) {
    return make_uint2((uint32_t)x0 | ((uint32_t)y0 << 16), ((uint32_t)x1 | ((uint32_t)y1 << 16)) | BOX_BIAS);
}

// M2b: block 0 numbers the new uniques in ascending current-CC order (:111-124), creates them and copies their crops
// into the arena (first-seen instance, never updated); the other blocks run the expiry pass (:127-145; not on frame 0)
// as an unordered compaction of the active list.  The last block out publishes the scalars of the next frame.
#define MU_THREADS 1024
#define MU_ITEMS 8
template <int THREADS>
__device__ __forceinline__ void match_update_body(int bid, int nblk, const CcFrame* frames, const int* __restrict__ counts, int f,
               const int* act_in, const uint2* box_in, int* act_out, uint2* box_out,
               int* scal, unsigned long long* scal64,
               int* u_min_x, int* u_max_x, int* u_min_y, int* u_max_y, int* u_size, int* u_last,
               int* u_first_frame, int* u_first_label, unsigned long long* u_crop_off, int* newlist,
               int MU, int MA, unsigned long long AW, int max_gap, int MP, int MI) {
    __shared__ int sm[33];
    const CcFrame fr = frames[f];
    const int n_kept = counts[f * 4 + 2];
    const int img_idx = scal[2], n_uniq0 = scal[0], n_act0 = scal[1];
    const int tid = threadIdx.x;
    bool over = false;
    if (bid == 0) {
        // every thread owns MU_ITEMS consecutive current CCs per round: all their loads are in flight together and
        // one block scan per round ranks the new ones in ascending order
        int n_new = 0;
        for (int c0 = 0; c0 < n_kept; c0 += THREADS * MU_ITEMS) {
            const int cb = c0 + tid * MU_ITEMS;
            unsigned newmask = 0;
#pragma unroll
            for (int k = 0; k < MU_ITEMS; ++k)
                if (cb + k < n_kept && fr.match_unique[cb + k] == MATCH_NONE) newmask |= 1u << k;
            int tot;
            int rank = n_new + block_excl_scan(__popc(newmask), sm, &tot);
            while (newmask) {
                const int k = __ffs(newmask) - 1; newmask &= newmask - 1;
                const int c = cb + k;
                const int l = fr.kept_label[c] - 1;
                const int x0 = fr.t_min_x[l], x1 = fr.t_max_x[l], y0 = fr.t_min_y[l], y1 = fr.t_max_y[l];
                const int words = ((x1 >> 5) - (x0 >> 5) + 1) * (y1 - y0 + 1);
                const int u = n_uniq0 + rank;
                const int ai = atomicAdd(&scal[7], 1);
                const unsigned long long off = atomicAdd(&scal64[1], (unsigned long long)words);
                if (u < MU && ai < MA && off + words <= AW) {
                    u_min_x[u] = x0; u_max_x[u] = x1; u_min_y[u] = y0; u_max_y[u] = y1;
                    u_size[u] = fr.t_count[l]; u_last[u] = img_idx; u_first_frame[u] = img_idx; u_first_label[u] = l + 1;
                    u_crop_off[u] = off;
                    act_out[ai] = u; box_out[ai] = pack_box(x0, x1, y0, y1);
                    fr.match_unique[c] = u;
                    newlist[rank] = c;
                } else { over = true; if (rank < MA) newlist[rank] = -1; }
                ++rank;
            }
            n_new += tot;
        }
        if (tid == 0) {
            scal[8] = n_new; scal[9] = min(n_new, MA);                   // [9]: crop copies for k_match_copy
            if (scal[4] > MP || scal[5] > MI) atomicOr(&scal[3], 16);     // candidate work lists overflowed
        }
    } else {
        const int lane = tid & 31;
        for (int i0 = (bid - 1) * THREADS; i0 < n_act0; i0 += (nblk - 1) * THREADS) {
            const int i = i0 + tid;
            const int u = (i < n_act0) ? act_in[i] : -1;
            const bool keep = (u >= 0) && (img_idx == 0 || img_idx - u_last[u] < max_gap);
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (bal == 0) continue;
            int base = 0;
            if (lane == 0) base = atomicAdd(&scal[7], __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) {
                const int ai = base + __popc(bal & ((1u << lane) - 1u));
                if (ai < MA) { act_out[ai] = u; box_out[ai] = box_in[i]; } else over = true;
            }
        }
    }
    if (__syncthreads_or(over) && tid == 0) atomicOr(&scal[3], 8);
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&scal[6], 1) == (int)nblk - 1) {              // last block out
            __threadfence();
            const int n_new = ((volatile int*)scal)[8], n_act = ((volatile int*)scal)[7];
            scal[0] = min(n_uniq0 + n_new, MU);
            scal[1] = min(n_act, MA);
            scal[2] = img_idx + 1;
            scal[4] = 0; scal[5] = 0; scal[6] = 0; scal[7] = 0; scal[8] = 0;   // empty work lists for the next frame
        }
    }
}
__global__ void __launch_bounds__(MU_THREADS)
k_match_update(const CcFrame* frames, const int* __restrict__ counts, int f, const int* act_in, const uint2* box_in, int* act_out, uint2* box_out,
               int* scal, unsigned long long* scal64, int* u_min_x, int* u_max_x, int* u_min_y, int* u_max_y, int* u_size, int* u_last,
               int* u_first_frame, int* u_first_label, unsigned long long* u_crop_off, int* newlist, int MU, int MA, unsigned long long AW,
               int max_gap, int MP, int MI) {
    match_update_body<MU_THREADS>(blockIdx.x, gridDim.x, frames, counts, f, act_in, box_in, act_out, box_out, scal, scal64, u_min_x, u_max_x, u_min_y,
                                  u_max_y, u_size, u_last, u_first_frame, u_first_label, u_crop_off, newlist, MU, MA, AW, max_gap, MP, MI);
}

// M2c: the crops of the new uniques go into the arena (first-seen instance, never updated): warp per new unique
__device__ __forceinline__ void match_copy_body(int bid, int nblk, const CcFrame* frames, int f, const int* scal, const int* newlist,
                             const unsigned long long* u_crop_off, uint32_t* arena) {
    const CcFrame fr = frames[f];
    const int n = scal[9], lane = threadIdx.x & 31;
    for (int i = (bid * blockDim.x + threadIdx.x) >> 5; i < n; i += (nblk * blockDim.x) >> 5) {
        const int c = newlist[i];
        if (c < 0) continue;
        const int l = fr.kept_label[c] - 1;
        const int words = ((fr.t_max_x[l] >> 5) - (fr.t_min_x[l] >> 5) + 1) * (fr.t_max_y[l] - fr.t_min_y[l] + 1);
        const uint32_t* src = fr.crops + fr.kept_crop_off[c];
        uint32_t* dst = arena + u_crop_off[fr.match_unique[c]];
        for (int k = lane; k < words; k += 32) dst[k] = src[k];
    }
}
__global__ void k_match_copy(const CcFrame* frames, int f, const int* scal, const int* newlist, const unsigned long long* u_crop_off,
                             uint32_t* arena) {
    match_copy_body(blockIdx.x, gridDim.x, frames, f, scal, newlist, u_crop_off, arena);
}

// All six phases of `n` consecutive frames in ONE cooperative launch: grid-wide barriers replace the kernel boundaries (6 launches
// and their ~5 us dependency gaps per frame -> 5 grid syncs).  On the headline workload (a handful of CCs per frame) the matching
// of an 8-frame batch drops from 48 launches to one; dense frames keep the same parallelism (the bodies are grid-strided).
struct MatchArgs {
    const CcFrame* frames; const int* counts; int first, n, cur;
    int* act[2]; uint2* box[2];
    int* scal; unsigned long long* scal64;
    int *u_min_x, *u_max_x, *u_min_y, *u_max_y, *u_size, *u_last, *u_first_frame, *u_first_label;
    unsigned long long* u_crop_off; uint32_t* arena; int* newlist;
    int2* pair_cu; int* pair_m; int2* items;
    int MU, MA, MP, MI, max_gap; unsigned long long AW; double min_recall, min_precision;
};
__global__ void __launch_bounds__(MP_THREADS, 4) k_match_fused(const MatchArgs a) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int bid = blockIdx.x, nblk = gridDim.x;
    int cur = a.cur;
    for (int f = a.first; f < a.first + a.n; ++f, cur ^= 1) {
        match_pairs_body(bid, nblk, a.frames, a.counts, f, a.act[cur], a.box[cur], a.scal, a.pair_cu, a.pair_m, a.items, a.MP, a.MI, a.scal64);
        grid.sync();
        match_overlap_body(bid, nblk, a.frames, f, a.scal, a.u_min_x, a.u_max_x, a.u_min_y, a.u_max_y, a.u_crop_off, a.arena, a.pair_cu, a.pair_m,
                           a.items, a.MP, a.MI);
        grid.sync();
        match_select_body(bid, nblk, a.frames, f, a.scal, a.u_size, a.pair_cu, a.pair_m, a.MP, a.min_recall, a.min_precision);
        grid.sync();
        match_refresh_body(bid, nblk, a.frames, a.counts, f, a.scal, a.u_last);
        grid.sync();
        match_update_body<MP_THREADS>(bid, nblk, a.frames, a.counts, f, a.act[cur], a.box[cur], a.act[cur ^ 1], a.box[cur ^ 1], a.scal, a.scal64,
                                      a.u_min_x, a.u_max_x, a.u_min_y, a.u_max_y, a.u_size, a.u_last, a.u_first_frame, a.u_first_label, a.u_crop_off,
                                      a.newlist, a.MU, a.MA, a.AW, a.max_gap, a.MP, a.MI);
        grid.sync();
        // the crop copies only have to land before the NEXT frame's overlap phase (one grid sync away)
        match_copy_body(bid, nblk, a.frames, f, a.scal, a.newlist, a.u_crop_off, a.arena);
    }
}

// packed boxes of the current active list (after an import)
__global__ void k_rebuild_boxes(const int* __restrict__ act, const int* __restrict__ scal, const int* u_min_x, const int* u_max_x,
                                const int* u_min_y, const int* u_max_y, uint2* __restrict__ box) {
    const int n = scal[1];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int u = act[i];
        box[i] = pack_box(u_min_x[u], u_max_x[u], u_min_y[u], u_max_y[u]);
    }
}

// ------------------------------------------------------------------------------------------------
// result rows: unique_idx, raw_label, min_x, max_x, min_y, max_y, size, crop_offset
__device__ __forceinline__ void write_row(const CcFrame& fr, int c, int* row) {
    int l = fr.kept_label[c] - 1;
    row[0] = fr.match_unique[c]; row[1] = l + 1; row[2] = fr.t_min_x[l]; row[3] = fr.t_max_x[l];
    row[4] = fr.t_min_y[l]; row[5] = fr.t_max_y[l]; row[6] = fr.t_count[l]; row[7] = (int)fr.kept_crop_off[c];
}
__global__ void k_rows_one(const CcFrame* __restrict__ frames, int f, int n, int* __restrict__ rows) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) write_row(frames[f], c, rows + (size_t)c * 8);
}
__global__ void k_row_offsets(const int* __restrict__ counts, int B, int* __restrict__ offs) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int s = 0;
        for (int f = 0; f < B; ++f) { offs[f] = s; s += counts[f * 4 + 2]; }
        offs[B] = s;
    }
}
__global__ void k_rows_all(const CcFrame* __restrict__ frames, const int* __restrict__ counts, const int* __restrict__ offs,
                           int cap, int* __restrict__ rows) {
    const int f = blockIdx.y;
    int n = counts[f * 4 + 2];
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    int r = offs[f] + c;
    if (r < cap) write_row(frames[f], c, rows + (size_t)r * 8);
}

// ------------------------------------------------------------------------------------------------
// active-set export / import (frame-shard hand-off)
__device__ __forceinline__ int crop_words_of(int mnx, int mxx, int mny, int mxy) {
    return ((mxx >> 5) - (mnx >> 5) + 1) * (mxy - mny + 1);
}
__global__ void k_export_sizes(const int* __restrict__ act, const int* __restrict__ scal, const int* u_min_x, const int* u_max_x,
                               const int* u_min_y, const int* u_max_y, unsigned long long* __restrict__ out) {
    __shared__ unsigned long long sm[32];
    int n = scal[1];
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int u = act[i];
        s += (unsigned long long)crop_words_of(u_min_x[u], u_max_x[u], u_min_y[u], u_max_y[u]);
    }
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i];
        out[0] = (unsigned long long)n; out[1] = t;
    }
}
// one block: meta rows + crop offsets (running), then warp-per-unique crop copy
__global__ void k_export(const int* __restrict__ act, const int* __restrict__ scal, const int* u_min_x, const int* u_max_x,
                         const int* u_min_y, const int* u_max_y, const int* u_size, const int* u_last, const int* u_ff,
                         const int* u_fl, const unsigned long long* u_crop_off, const uint32_t* __restrict__ arena,
                         int* __restrict__ meta, uint32_t* __restrict__ crops, unsigned long long* __restrict__ tmp_off) {
    __shared__ int sm[33];
    int n = scal[1];
    unsigned long long carry = 0;
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
        int i = i0 + threadIdx.x;
        int words = 0, u = 0;
        if (i < n) { u = act[i]; words = crop_words_of(u_min_x[u], u_max_x[u], u_min_y[u], u_max_y[u]); }
        int tot;
        int ex = block_excl_scan(words, sm, &tot);
        __syncthreads();
        if (i < n) {
            int* m = meta + (size_t)i * 10;
            m[0] = u; m[1] = u_min_x[u]; m[2] = u_max_x[u]; m[3] = u_min_y[u]; m[4] = u_max_y[u]; m[5] = u_size[u];
            m[6] = u_last[u]; m[7] = u_ff[u]; m[8] = u_fl[u]; m[9] = words;
            tmp_off[i] = carry + (unsigned long long)ex;
        }
        carry += (unsigned long long)tot;
    }
    __syncthreads();
    for (int i = threadIdx.x >> 5; i < n; i += (blockDim.x >> 5)) {
        int u = act[i];
        int words = meta[(size_t)i * 10 + 9];
        const uint32_t* src = arena + u_crop_off[u];
        uint32_t* dst = crops + tmp_off[i];
        for (int k = threadIdx.x & 31; k < words; k += 32) dst[k] = src[k];
    }
}
__global__ void k_import(int n, const int* __restrict__ meta, const uint32_t* __restrict__ crops, int* act, int* u_min_x,
                         int* u_max_x, int* u_min_y, int* u_max_y, int* u_size, int* u_last, int* u_ff, int* u_fl,
                         unsigned long long* u_crop_off, uint32_t* arena, unsigned long long* tmp_off, int MU) {
    __shared__ int sm[33];
    unsigned long long carry = 0;
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
        int i = i0 + threadIdx.x;
        int words = (i < n) ? meta[(size_t)i * 10 + 9] : 0;
        int tot;
        int ex = block_excl_scan(words, sm, &tot);
        __syncthreads();
        if (i < n) {
            const int* m = meta + (size_t)i * 10;
            int u = m[0];
            if (u >= 0 && u < MU) {
                u_min_x[u] = m[1]; u_max_x[u] = m[2]; u_min_y[u] = m[3]; u_max_y[u] = m[4]; u_size[u] = m[5];
                u_last[u] = m[6]; u_ff[u] = m[7]; u_fl[u] = m[8];
                u_crop_off[u] = carry + (unsigned long long)ex;
            }
            act[i] = u;
            tmp_off[i] = carry + (unsigned long long)ex;
        }
        carry += (unsigned long long)tot;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (int)min(carry, (unsigned long long)0x7fffffff); i += blockDim.x) arena[i] = crops[i];
}

// ------------------------------------------------------------------------------------------------
// Fully asynchronous hand-off (no host read-back of sizes): the active set travels in ONE fixed-capacity buffer
//   buf[0..16)  header: n_active, crop_words, n_unique, img_idx, tempo_count lo/hi, flags (1 = did not fit), 0...
//   buf[16 ..)  meta [n_active][10] (same rows as k_export), then the crops, concatenated in meta order
#define XHDR 16
__global__ void k_export_dev_meta(const int* __restrict__ act, const int* __restrict__ scal, const unsigned long long* __restrict__ scal64,
                                  const int* u_min_x, const int* u_max_x, const int* u_min_y, const int* u_max_y, const int* u_size,
                                  const int* u_last, const int* u_ff, const int* u_fl, int* __restrict__ buf, long long cap,
                                  unsigned long long* __restrict__ tmp_off) {
    __shared__ int sm[33];
    const int n = scal[1];
    const bool meta_fits = (long long)XHDR + (long long)n * 10 <= cap;
    unsigned long long carry = 0;
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
        int i = i0 + threadIdx.x;
        int words = 0, u = 0;
        if (i < n) { u = act[i]; words = crop_words_of(u_min_x[u], u_max_x[u], u_min_y[u], u_max_y[u]); }
        int tot;
        int ex = block_excl_scan(words, sm, &tot);
        __syncthreads();
        if (i < n) {
            if (meta_fits) {
                int* m = buf + XHDR + (size_t)i * 10;
                m[0] = u; m[1] = u_min_x[u]; m[2] = u_max_x[u]; m[3] = u_min_y[u]; m[4] = u_max_y[u]; m[5] = u_size[u];
                m[6] = u_last[u]; m[7] = u_ff[u]; m[8] = u_fl[u]; m[9] = words;
            }
            tmp_off[i] = carry + (unsigned long long)ex;
        }
        carry += (unsigned long long)tot;
    }
    if (threadIdx.x == 0) {
        const bool fits = (long long)XHDR + (long long)n * 10 + (long long)carry <= cap;
        buf[0] = n; buf[1] = (int)carry; buf[2] = scal[0]; buf[3] = scal[2];
        buf[4] = (int)(scal64[0] & 0xffffffffull); buf[5] = (int)(scal64[0] >> 32);
        buf[6] = (fits ? 0 : 1) | (scal[3] ? 2 : 0);                   // 1: hand-off buffer too small, 2: sender already overflowed
        for (int k = 7; k < XHDR; ++k) buf[k] = 0;
    }
}
__global__ void k_export_dev_crops(const int* __restrict__ act, const int* __restrict__ scal, const int* u_min_x, const int* u_max_x,
                                   const int* u_min_y, const int* u_max_y, const unsigned long long* __restrict__ u_crop_off,
                                   const uint32_t* __restrict__ arena, int* __restrict__ buf, long long cap,
                                   const unsigned long long* __restrict__ tmp_off) {
    const int n = scal[1];
    const long long base = (long long)XHDR + (long long)n * 10;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int u = act[i];
        const int words = crop_words_of(u_min_x[u], u_max_x[u], u_min_y[u], u_max_y[u]);
        const long long dst0 = base + (long long)tmp_off[i];
        if (dst0 + words > cap) continue;                               // flagged in the header
        const uint32_t* src = arena + u_crop_off[u];
        uint32_t* dst = (uint32_t*)buf + dst0;
        for (int k = threadIdx.x; k < words; k += blockDim.x) dst[k] = src[k];
    }
}
__global__ void k_import_dev_meta(const int* __restrict__ buf, int* act, int* scal, unsigned long long* scal64, int* u_min_x, int* u_max_x,
                                  int* u_min_y, int* u_max_y, int* u_size, int* u_last, int* u_ff, int* u_fl,
                                  unsigned long long* u_crop_off, int MU, int MA, unsigned long long AW) {
    __shared__ int sm[33];
    const int n = buf[0];
    const bool ok = buf[6] == 0 && n <= MA && buf[2] <= MU && (unsigned long long)(unsigned)buf[1] <= AW;
    const int* meta = buf + XHDR;
    unsigned long long carry = 0;
    for (int i0 = 0; ok && i0 < n; i0 += blockDim.x) {
        int i = i0 + threadIdx.x;
        int words = (i < n) ? meta[(size_t)i * 10 + 9] : 0;
        int tot;
        int ex = block_excl_scan(words, sm, &tot);
        __syncthreads();
        if (i < n) {
            const int* m = meta + (size_t)i * 10;
            int u = m[0];
            if (u >= 0 && u < MU) {
                u_min_x[u] = m[1]; u_max_x[u] = m[2]; u_min_y[u] = m[3]; u_max_y[u] = m[4]; u_size[u] = m[5];
                u_last[u] = m[6]; u_ff[u] = m[7]; u_fl[u] = m[8];
                u_crop_off[u] = carry + (unsigned long long)ex;
            }
            act[i] = u;
        }
        carry += (unsigned long long)tot;
    }
    if (threadIdx.x == 0) {
        scal[0] = buf[2]; scal[1] = ok ? n : 0; scal[2] = buf[3]; scal[3] = ok ? 0 : 32;   // 32: hand-off failed
        scal[4] = 0; scal[5] = 0;
        scal64[0] = ((unsigned long long)(unsigned)buf[5] << 32) | (unsigned long long)(unsigned)buf[4];
        scal64[1] = ok ? (unsigned long long)(unsigned)buf[1] : 0ull;
    }
}
__global__ void k_import_dev_crops(const int* __restrict__ buf, uint32_t* __restrict__ arena, unsigned long long AW) {
    if (buf[6] != 0) return;
    const long long base = (long long)XHDR + (long long)buf[0] * 10;
    const long long n = min((long long)(unsigned)buf[1], (long long)AW);
    const uint32_t* src = (const uint32_t*)buf + base;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) arena[i] = src[i];
}

// ================================================================================================
// Host side: C ABI
// ================================================================================================
static inline cudaStream_t S(void* s) { return (cudaStream_t)s; }
static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" int am_version(void) { return 100; }
extern "C" int am_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int am_words_per_row(int width) { return am_words_per_row_impl(width); }

extern "C" int am_pack_mask_u8(const uint8_t* d_mask, int width, int height, int batch, uint32_t* d_bits, void* stream) {
    if (!d_mask || !d_bits || width <= 0 || height <= 0 || batch <= 0) return AM_ERR_ARG;
    int WPR = am_words_per_row_impl(width);
    dim3 grid(am_div_up((long long)WPR * 32, 256), height, batch);
    k_pack_u8<<<grid, 256, 0, S(stream)>>>(d_mask, width, height, WPR, d_bits);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}
extern "C" int am_pack_mask_u8_exact(const uint8_t* d_mask, int width, int height, int batch, uint32_t* d_bits, int* d_flags, void* stream) {
    if (!d_mask || !d_bits || !d_flags || width <= 0 || height <= 0 || batch <= 0) return AM_ERR_ARG;
    const int WPR = am_words_per_row_impl(width);
    AM_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int) * batch, S(stream)));
    dim3 grid(am_div_up(WPR * 32, 256), height, batch);
    k_pack_u8_exact<<<grid, 256, 0, S(stream)>>>(d_mask, width, height, WPR, d_bits, d_flags);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}
extern "C" int am_unpack_mask_u8(const uint32_t* d_bits, int width, int height, int batch, uint8_t* d_mask, void* stream) {
    if (!d_mask || !d_bits || width <= 0 || height <= 0 || batch <= 0) return AM_ERR_ARG;
    dim3 grid(am_div_up(width, 256), height, batch);
    k_unpack_u8<<<grid, 256, 0, S(stream)>>>(d_bits, width, height, am_words_per_row_impl(width), d_mask);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}

extern "C" am_cc_ctx* am_cc_create(int width, int height, int max_batch, int max_runs, int max_labels, int max_kept,
                                   int crop_words, int min_pixels) {
    (void)max_runs;                                      // run / strip-component slots are sized for the worst case
    if (width <= 0 || height <= 0 || width > 65535 || height > 32767 || max_batch <= 0) return nullptr;
    am_cc_ctx* c = new am_cc_ctx();
    c->W = width; c->H = height; c->WPR = am_words_per_row_impl(width); c->B = max_batch;
    long long P = (long long)width * height;
    // strip height: the largest power of two whose worst-case run count (every other pixel) fits the shared-memory
    // union-find of k_strip_label
    const int per_row = (width + 1) / 2;
    int R = 32;
    while (R > 1 && (long long)R * per_row > STRIP_MAX_RUNS) R >>= 1;
    c->R = R; c->NS = (height + R - 1) / R; c->CAP = R * per_row;
    c->strip_smem = (2 * R * c->WPR + c->CAP + 5 * STRIP_CAPS) * 4;
    if (c->strip_smem > 200 * 1024) { fprintf(stderr, "[accessmath_b200] am_cc_create: width %d too large\n", width); delete c; return nullptr; }
    c->ML = max_labels > 0 ? max_labels : (int)(P / 2 + 64);
    c->MK = max_kept > 0 ? max_kept : (int)(P / (min_pixels > 1 ? min_pixels : 1) + 64);      // a kept CC has >= min_pixels pixels
    if (c->MK > c->ML) c->MK = c->ML;
    c->CW = crop_words > 0 ? crop_words : (int)(4 * (long long)c->WPR * height + 1024);
    c->min_pixels = min_pixels;
    const size_t slots = (size_t)c->NS * c->CAP;
    size_t per = 0;
    auto add = [&](size_t bytes) { size_t o = per; per += align_up(bytes); return o; };
    size_t o_wf = add((size_t)height * c->WPR * 4), o_rc = add(slots * 4);
    size_t o_c[7]; for (int i = 0; i < 7; ++i) o_c[i] = add(slots * 4);
    size_t o_sn = add((size_t)c->NS * sizeof(int2));
    size_t o_t[5]; for (int i = 0; i < 5; ++i) o_t[i] = add((size_t)c->ML * 4);
    size_t o_lco = add((size_t)c->ML * 4);
    size_t o_kl = add((size_t)c->MK * 4), o_kco = add((size_t)c->MK * 4), o_mu = add((size_t)c->MK * 4);
    size_t o_cr = add((size_t)(c->CW + 4) * 4);
    size_t head = align_up(sizeof(CcFrame) * max_batch) + align_up((size_t)max_batch * 16) + 256;
    size_t total = head + per * max_batch;
    if (cudaMalloc(&c->slab, total) != cudaSuccess) {
        fprintf(stderr, "[accessmath_b200] am_cc_create: cudaMalloc(%zu) failed\n", total);
        cudaGetLastError();
        delete c; return nullptr;
    }
    char* base = (char*)c->slab;
    c->d_frames = (CcFrame*)base;
    c->d_counts = (int*)(base + align_up(sizeof(CcFrame) * max_batch));
    c->d_status = (int*)(base + align_up(sizeof(CcFrame) * max_batch) + align_up((size_t)max_batch * 16));
    c->h_frames = new CcFrame[max_batch];
    for (int f = 0; f < max_batch; ++f) {
        char* p = base + head + per * f;
        CcFrame& fr = c->h_frames[f];
        fr.wfirst = (uint32_t*)(p + o_wf); fr.run_comp = (uint32_t*)(p + o_rc);
        fr.c_min_x = (int*)(p + o_c[0]); fr.c_max_x = (int*)(p + o_c[1]); fr.c_min_y = (int*)(p + o_c[2]); fr.c_max_y = (int*)(p + o_c[3]);
        fr.c_count = (int*)(p + o_c[4]); fr.c_link = (int*)(p + o_c[5]); fr.c_label = (int*)(p + o_c[6]);
        fr.strip_n = (int2*)(p + o_sn);
        fr.t_min_y = (int*)(p + o_t[0]); fr.t_max_y = (int*)(p + o_t[1]); fr.t_min_x = (int*)(p + o_t[2]);
        fr.t_max_x = (int*)(p + o_t[3]); fr.t_count = (int*)(p + o_t[4]); fr.lab_crop_off = (uint32_t*)(p + o_lco);
        fr.kept_label = (int*)(p + o_kl); fr.kept_crop_off = (uint32_t*)(p + o_kco); fr.match_unique = (int*)(p + o_mu);
        fr.crops = (uint32_t*)(p + o_cr);
    }
    cudaMemcpy(c->d_frames, c->h_frames, sizeof(CcFrame) * max_batch, cudaMemcpyHostToDevice);
    cudaMemset(c->d_counts, 0, (size_t)max_batch * 16);
    cudaMemset(c->d_status, 0, 4);
    cudaMallocHost(&c->h_counts, (size_t)max_batch * 16 + 16);
    cudaFuncSetAttribute(k_strip_label, cudaFuncAttributeMaxDynamicSharedMemorySize, c->strip_smem);
    if ((size_t)(c->NS + 1) * 4 > 48 * 1024) cudaFuncSetAttribute(k_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (c->NS + 1) * 4);
    return c;
}
extern "C" void am_cc_destroy(am_cc_ctx* c) {
    if (!c) return;
    cudaFree(c->slab); cudaFreeHost(c->h_counts); delete[] c->h_frames; delete c;
}

extern "C" int am_cc_label_batch(am_cc_ctx* c, const uint32_t* d_bits, int batch, int32_t* d_labels, void* stream) {
    if (!c || !d_bits || batch <= 0 || batch > c->B) return AM_ERR_ARG;
    cudaStream_t st = S(stream);
    const int H = c->H, W = c->W, WPR = c->WPR;
    k_strip_label<<<dim3(c->NS, batch), STRIP_THREADS, c->strip_smem, st>>>(d_bits, c->d_frames, W, H, WPR, c->R, c->CAP);
    k_resolve<<<batch, RESOLVE_THREADS, (c->NS + 1) * 4, st>>>(d_bits, c->d_frames, c->d_counts, H, WPR, c->R, c->NS, c->CAP, c->ML, c->MK,
                                                             c->CW, c->min_pixels, c->d_status);
    const int nwords = H * WPR;
    k_crop_fill<<<dim3(nwords >= 128 * 256 ? 128 : am_div_up(nwords, 256), batch), 256, 0, st>>>(d_bits, c->d_frames, H, WPR);
    if (d_labels) {
        const int ngroups = H * ((W + 3) / 4);
        k_label_image<<<dim3(ngroups >= 256 * 256 ? 256 : am_div_up(ngroups, 256), batch), 256, 0, st>>>(d_bits, c->d_frames, W, H, WPR, d_labels);
    }
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}

extern "C" int am_cc_counts(am_cc_ctx* c, int batch, int* h_counts, void* stream) {
    if (!c || !h_counts || batch <= 0 || batch > c->B) return AM_ERR_ARG;
    AM_CUDA(cudaMemcpyAsync(c->h_counts, c->d_counts, (size_t)batch * 16, cudaMemcpyDeviceToHost, S(stream)));
    AM_CUDA(cudaMemcpyAsync(c->h_counts + batch * 4, c->d_status, 4, cudaMemcpyDeviceToHost, S(stream)));
    AM_CUDA(cudaStreamSynchronize(S(stream)));
    memcpy(h_counts, c->h_counts, (size_t)batch * 16);
    if (c->h_counts[batch * 4] != 0) {
        fprintf(stderr, "[accessmath_b200] CC capacity exceeded (flags 0x%x: 1=runs 2=labels 4=kept/crops)\n", c->h_counts[batch * 4]);
        return AM_ERR_CAPACITY;
    }
    return AM_OK;
}

extern "C" int am_cc_read_label_table(am_cc_ctx* c, int frame, int n, int* h_min_y, int* h_max_y, int* h_min_x, int* h_max_x,
                                      int* h_count, void* stream) {
    if (!c || frame < 0 || frame >= c->B || n < 0 || n > c->ML) return AM_ERR_ARG;
    const CcFrame& fr = c->h_frames[frame];
    size_t b = (size_t)n * 4;
    if (n) {
        AM_CUDA(cudaMemcpyAsync(h_min_y, fr.t_min_y, b, cudaMemcpyDeviceToHost, S(stream)));
        AM_CUDA(cudaMemcpyAsync(h_max_y, fr.t_max_y, b, cudaMemcpyDeviceToHost, S(stream)));
        AM_CUDA(cudaMemcpyAsync(h_min_x, fr.t_min_x, b, cudaMemcpyDeviceToHost, S(stream)));
        AM_CUDA(cudaMemcpyAsync(h_max_x, fr.t_max_x, b, cudaMemcpyDeviceToHost, S(stream)));
        AM_CUDA(cudaMemcpyAsync(h_count, fr.t_count, b, cudaMemcpyDeviceToHost, S(stream)));
    }
    AM_CUDA(cudaStreamSynchronize(S(stream)));
    return AM_OK;
}

extern "C" int am_cc_read_kept(am_cc_ctx* c, int frame, int n_kept, int* h_rows, void* stream) {
    if (!c || frame < 0 || frame >= c->B || n_kept < 0 || n_kept > c->MK) return AM_ERR_ARG;
    if (n_kept == 0) return AM_OK;
    int* d_rows = nullptr;
    AM_CUDA(cudaMallocAsync(&d_rows, (size_t)n_kept * 32, S(stream)));
    k_rows_one<<<am_div_up(n_kept, 256), 256, 0, S(stream)>>>(c->d_frames, frame, n_kept, d_rows);
    AM_CUDA(cudaMemcpyAsync(h_rows, d_rows, (size_t)n_kept * 32, cudaMemcpyDeviceToHost, S(stream)));
    AM_CUDA(cudaFreeAsync(d_rows, S(stream)));
    AM_CUDA(cudaStreamSynchronize(S(stream)));
    return AM_OK;
}

extern "C" int am_cc_read_crops(am_cc_ctx* c, int frame, int crop_words, uint32_t* h_crops, void* stream) {
    if (!c || frame < 0 || frame >= c->B || crop_words < 0 || crop_words > c->CW) return AM_ERR_ARG;
    if (crop_words) AM_CUDA(cudaMemcpyAsync(h_crops, c->h_frames[frame].crops, (size_t)crop_words * 4, cudaMemcpyDeviceToHost, S(stream)));
    AM_CUDA(cudaStreamSynchronize(S(stream)));
    return AM_OK;
}

extern "C" int am_cc_pack_rows(am_cc_ctx* c, int batch, int* d_rows, int row_capacity, int* d_row_offsets, void* stream) {
    if (!c || !d_rows || !d_row_offsets || batch <= 0 || batch > c->B) return AM_ERR_ARG;
    k_row_offsets<<<1, 32, 0, S(stream)>>>(c->d_counts, batch, d_row_offsets);
    k_rows_all<<<dim3(am_div_up(c->MK, 256), batch), 256, 0, S(stream)>>>(c->d_frames, c->d_counts, d_row_offsets, row_capacity, d_rows);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}

// ---- legacy host-pointer operator -----------------------------------------------------------
extern "C" int CC_AgeBoundaries(int* labels, float* ages, int width, int height, int count_labels, int* out_mins_y,
                                int* out_maxs_y, int* out_mins_x, int* out_maxs_x, int* out_counts, float* output_age) {
    if (count_labels <= 0) return 0;
    if (!labels || width <= 0 || height <= 0) return AM_ERR_ARG;
    size_t P = (size_t)width * height, n = (size_t)count_labels;
    int32_t* d_lab = nullptr; float* d_age = nullptr; int* d_tab = nullptr;
    int* outs[6] = {out_mins_y, out_maxs_y, out_mins_x, out_maxs_x, out_counts, (int*)output_age};
    const int src[6] = {0, 1, 2, 3, 4, 6};
    int rc = AM_ERR_CUDA;
    do {                                                   // single exit: the temporaries are freed on every path
        if (cudaMalloc(&d_lab, P * 4) != cudaSuccess) break;
        if (ages && cudaMalloc(&d_age, P * 4) != cudaSuccess) break;
        if (cudaMalloc(&d_tab, (n * 10 + 4) * 4) != cudaSuccess) break;
        if (cudaMemcpy(d_lab, labels, P * 4, cudaMemcpyHostToDevice) != cudaSuccess) break;
        if (ages && cudaMemcpy(d_age, ages, P * 4, cudaMemcpyHostToDevice) != cudaSuccess) break;
        if (ageb_run(d_lab, d_age, width, height, count_labels, d_tab, 0) != AM_OK) break;
        bool ok = true;
        for (int i = 0; i < 6 && ok; ++i)
            ok = cudaMemcpy(outs[i], d_tab + (size_t)src[i] * n, n * 4, cudaMemcpyDeviceToHost) == cudaSuccess;
        if (ok) rc = 0;
    } while (0);
    if (rc) { cudaError_t e = cudaGetLastError(); fprintf(stderr, "[accessmath_b200] CC_AgeBoundaries: CUDA error %d (%s)\n", (int)e, cudaGetErrorString(e)); }
    cudaFree(d_lab); cudaFree(d_age); cudaFree(d_tab);
    return rc;
}
// The same operator on DEVICE pointers (labels int32 [H][W], ages fp32 [H][W] or NULL = all zero; outputs 6 x n, d_scratch >= 10 n + 4
// ints), asynchronous on `stream`: what Labeler uses when the label image is already resident (no PCIe round trip of 8 P bytes).
extern "C" int am_cc_age_boundaries_dev(const int* d_labels, const float* d_ages, int width, int height, int count_labels, int* d_out6,
                                        int* d_scratch, void* stream) {
    if (count_labels <= 0) return AM_OK;
    if (!d_labels || !d_out6 || !d_scratch || width <= 0 || height <= 0) return AM_ERR_ARG;
    const size_t n = (size_t)count_labels;
    int rc = ageb_run(d_labels, d_ages, width, height, count_labels, d_scratch, S(stream));
    if (rc) return rc;
    const int src[6] = {0, 1, 2, 3, 4, 6};
    for (int i = 0; i < 6; ++i)
        AM_CUDA(cudaMemcpyAsync(d_out6 + (size_t)i * n, d_scratch + (size_t)src[i] * n, n * 4, cudaMemcpyDeviceToDevice, S(stream)));
    return AM_OK;
}

// ---- estimator ------------------------------------------------------------------------------
extern "C" am_estimator* am_est_create(int width, int height, double min_recall, double min_precision, int max_gap,
                                       int max_uniques, int max_active, long long arena_words) {
    if (width <= 0 || height <= 0 || width > 32767 || height > 32767) return nullptr;      // packed 15-bit boxes
    am_estimator* e = new am_estimator();
    e->W = width; e->H = height; e->max_gap = max_gap; e->min_recall = min_recall; e->min_precision = min_precision;
    e->MU = max_uniques > 0 ? max_uniques : (1 << 20);
    e->MA = max_active > 0 ? max_active : (1 << 18);
    if (e->MA > e->MU) e->MA = e->MU;
    unsigned long long aw = arena_words > 0 ? (unsigned long long)arena_words : (64ull << 20);
    size_t tot = 0;
    auto add = [&](size_t bytes) { size_t o = tot; tot += align_up(bytes); return o; };
    size_t o_i[8]; for (int i = 0; i < 8; ++i) o_i[i] = add((size_t)e->MU * 4);
    size_t o_off = add((size_t)e->MU * 8);
    size_t o_a0 = add((size_t)e->MA * 4), o_a1 = add((size_t)e->MA * 4);
    size_t o_b0 = add((size_t)e->MA * 8), o_b1 = add((size_t)e->MA * 8), o_nl = add((size_t)e->MA * 4);
    size_t o_tmp = add((size_t)e->MA * 8);
    size_t o_sc = add(64), o_sc64 = add(64);
    e->MP = 1 << 20; e->MI = 1 << 21;
    size_t o_pcu = add((size_t)e->MP * 8), o_pm = add((size_t)e->MP * 4), o_it = add((size_t)e->MI * 8);
    size_t o_ar = add((size_t)aw * 4);
    if (cudaMalloc(&e->slab, tot) != cudaSuccess) {
        fprintf(stderr, "[accessmath_b200] am_est_create: cudaMalloc(%zu) failed\n", tot);
        delete e; return nullptr;
    }
    char* b = (char*)e->slab;
    e->u_min_x = (int*)(b + o_i[0]); e->u_max_x = (int*)(b + o_i[1]); e->u_min_y = (int*)(b + o_i[2]); e->u_max_y = (int*)(b + o_i[3]);
    e->u_size = (int*)(b + o_i[4]); e->u_last = (int*)(b + o_i[5]); e->u_first_frame = (int*)(b + o_i[6]); e->u_first_label = (int*)(b + o_i[7]);
    e->u_crop_off = (unsigned long long*)(b + o_off);
    e->act[0] = (int*)(b + o_a0); e->act[1] = (int*)(b + o_a1); e->cur = 0;
    e->box[0] = (uint2*)(b + o_b0); e->box[1] = (uint2*)(b + o_b1); e->newlist = (int*)(b + o_nl);
    e->d_scal = (int*)(b + o_sc); e->d_scal64 = (unsigned long long*)(b + o_sc64);
    e->arena = (uint32_t*)(b + o_ar);
    e->pair_cu = (int2*)(b + o_pcu); e->pair_m = (int*)(b + o_pm); e->items = (int2*)(b + o_it);
    e->AW = aw; e->tmp_off = (unsigned long long*)(b + o_tmp);
    cudaMemset(e->d_scal, 0, 64); cudaMemset(e->d_scal64, 0, 64);
    cudaMallocHost(&e->h_scal, 64); cudaMallocHost(&e->h_scal64, 64);
    return e;
}
extern "C" void am_est_destroy(am_estimator* e) {
    if (!e) return;
    cudaFree(e->slab); cudaFreeHost(e->h_scal); cudaFreeHost(e->h_scal64); delete e;
}
static inline unsigned long long* est_tmp(am_estimator* e) { return e->tmp_off; }

extern "C" int am_est_add_frames(am_estimator* e, am_cc_ctx* c, int first, int n, void* stream) {
    if (!e || !c || first < 0 || n <= 0 || first + n > c->B) return AM_ERR_ARG;
    cudaStream_t st = S(stream);
    // default: ONE cooperative launch for the whole batch (k_match_fused); AM_B200_MATCH=multi keeps the six launches per frame
    // the cooperative-launch configuration belongs to the estimator (= to the device it was created on), not to the process
    if (e->fused_mode < 0) {
        const char* env = getenv("AM_B200_MATCH");
        int dev = 0, coop = 0, sms = 0, occ = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_match_fused, MP_THREADS, 0);
        const char* genv = getenv("AM_B200_MATCH_CTAS_PER_SM");     // tuning knob (default 2): grid syncs get dearer with more CTAs
        int per_sm = genv ? atoi(genv) : 2;
        if (per_sm < 1) per_sm = 1;
        if (per_sm > occ) per_sm = occ;
        e->fused_grid = sms * per_sm;
        const char* gabs = getenv("AM_B200_MATCH_GRID");                // tuning knob: absolute CTA count of the cooperative launch
        if (gabs && atoi(gabs) >= 2 && atoi(gabs) <= sms * occ) e->fused_grid = atoi(gabs);
        e->fused_mode = (coop && e->fused_grid >= 2 && !(env && strcmp(env, "multi") == 0)) ? 1 : 0;
    }
    const int fused_mode = e->fused_mode, fused_grid = e->fused_grid;
    if (fused_mode == 1) {
        MatchArgs a;
        a.frames = c->d_frames; a.counts = c->d_counts; a.first = first; a.n = n; a.cur = e->cur;
        a.act[0] = e->act[0]; a.act[1] = e->act[1]; a.box[0] = e->box[0]; a.box[1] = e->box[1];
        a.scal = e->d_scal; a.scal64 = e->d_scal64;
        a.u_min_x = e->u_min_x; a.u_max_x = e->u_max_x; a.u_min_y = e->u_min_y; a.u_max_y = e->u_max_y; a.u_size = e->u_size;
        a.u_last = e->u_last; a.u_first_frame = e->u_first_frame; a.u_first_label = e->u_first_label; a.u_crop_off = e->u_crop_off;
        a.arena = e->arena; a.newlist = e->newlist; a.pair_cu = e->pair_cu; a.pair_m = e->pair_m; a.items = e->items;
        a.MU = e->MU; a.MA = e->MA; a.MP = e->MP; a.MI = e->MI; a.max_gap = e->max_gap; a.AW = e->AW;
        a.min_recall = e->min_recall; a.min_precision = e->min_precision;
        void* args[] = {&a};
        AM_CUDA(cudaLaunchCooperativeKernel((const void*)k_match_fused, dim3(fused_grid), dim3(MP_THREADS), args, 0, st));
        e->cur ^= (n & 1);
        return AM_OK;
    }
    for (int f = first; f < first + n; ++f) {
        int* a_in = e->act[e->cur]; int* a_out = e->act[e->cur ^ 1];
        uint2* b_in = e->box[e->cur]; uint2* b_out = e->box[e->cur ^ 1];
        k_match_pairs<<<148 * 4, MP_THREADS, 0, st>>>(c->d_frames, c->d_counts, f, a_in, b_in, e->d_scal, e->pair_cu, e->pair_m, e->items,
                                                    e->MP, e->MI, e->d_scal64);
        k_match_overlap<<<148 * 4, 256, 0, st>>>(c->d_frames, f, e->d_scal, e->u_min_x, e->u_max_x, e->u_min_y, e->u_max_y, e->u_crop_off,
                                                 e->arena, e->pair_cu, e->pair_m, e->items, e->MP, e->MI);
        k_match_select<<<148, 256, 0, st>>>(c->d_frames, f, e->d_scal, e->u_size, e->pair_cu, e->pair_m, e->MP, e->min_recall, e->min_precision);
        k_match_refresh<<<32, 256, 0, st>>>(c->d_frames, c->d_counts, f, e->d_scal, e->u_last);
        k_match_update<<<1 + 32, MU_THREADS, 0, st>>>(c->d_frames, c->d_counts, f, a_in, b_in, a_out, b_out, e->d_scal, e->d_scal64,
                                                     e->u_min_x, e->u_max_x, e->u_min_y, e->u_max_y, e->u_size, e->u_last, e->u_first_frame,
                                                     e->u_first_label, e->u_crop_off, e->newlist, e->MU, e->MA, e->AW, e->max_gap, e->MP, e->MI);
        k_match_copy<<<64, 256, 0, st>>>(c->d_frames, f, e->d_scal, e->newlist, e->u_crop_off, e->arena);
        e->cur ^= 1;
    }
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}

extern "C" int am_est_state(am_estimator* e, int* h_state, void* stream) {
    if (!e || !h_state) return AM_ERR_ARG;
    AM_CUDA(cudaMemcpyAsync(e->h_scal, e->d_scal, 16, cudaMemcpyDeviceToHost, S(stream)));
    AM_CUDA(cudaMemcpyAsync(e->h_scal64, e->d_scal64, 16, cudaMemcpyDeviceToHost, S(stream)));
    AM_CUDA(cudaStreamSynchronize(S(stream)));
    h_state[0] = e->h_scal[0]; h_state[1] = e->h_scal[1]; h_state[2] = e->h_scal[2]; h_state[3] = e->h_scal[3];
    h_state[4] = (int)(e->h_scal64[0] & 0xffffffffull); h_state[5] = (int)(e->h_scal64[0] >> 32);
    if (e->h_scal[3]) {
        fprintf(stderr, "[accessmath_b200] estimator capacity exceeded (flags 0x%x: 8=uniques/active/arena 16=candidate pairs 32=shard hand-off)\n", e->h_scal[3]);
        return AM_ERR_CAPACITY;
    }
    return AM_OK;
}

__global__ void k_uniq_rows(int first, int n, const int* ff, const int* fl, const int* mnx, const int* mxx, const int* mny,
                            const int* mxy, const int* sz, const int* last, int* rows) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int u = first + i; int* r = rows + (size_t)i * 8;
    r[0] = ff[u]; r[1] = fl[u]; r[2] = mnx[u]; r[3] = mxx[u]; r[4] = mny[u]; r[5] = mxy[u]; r[6] = sz[u]; r[7] = last[u];
}
extern "C" int am_est_read_uniques(am_estimator* e, int first, int n, int* h_rows, void* stream) {
    if (!e || first < 0 || n < 0 || first + n > e->MU) return AM_ERR_ARG;
    if (n == 0) return AM_OK;
    int* d_rows = nullptr;
    AM_CUDA(cudaMallocAsync(&d_rows, (size_t)n * 32, S(stream)));
    k_uniq_rows<<<am_div_up(n, 256), 256, 0, S(stream)>>>(first, n, e->u_first_frame, e->u_first_label, e->u_min_x, e->u_max_x,
                                                        e->u_min_y, e->u_max_y, e->u_size, e->u_last, d_rows);
    AM_CUDA(cudaMemcpyAsync(h_rows, d_rows, (size_t)n * 32, cudaMemcpyDeviceToHost, S(stream)));
    AM_CUDA(cudaFreeAsync(d_rows, S(stream)));
    AM_CUDA(cudaStreamSynchronize(S(stream)));
    return AM_OK;
}
extern "C" int am_est_read_unique_crop(am_estimator* e, int u, int words, uint32_t* h_crop, void* stream) {
    if (!e || u < 0 || u >= e->MU || words < 0) return AM_ERR_ARG;
    unsigned long long off = 0;
    AM_CUDA(cudaMemcpyAsync(&off, e->u_crop_off + u, 8, cudaMemcpyDeviceToHost, S(stream)));
    AM_CUDA(cudaStreamSynchronize(S(stream)));
    if (words) AM_CUDA(cudaMemcpyAsync(h_crop, e->arena + off, (size_t)words * 4, cudaMemcpyDeviceToHost, S(stream)));
    AM_CUDA(cudaStreamSynchronize(S(stream)));
    return AM_OK;
}

extern "C" int am_est_unique_view(am_estimator* e, am_unique_view* out, void* stream) {
    if (!e || !out) return AM_ERR_ARG;
    int n = 0;
    AM_CUDA(cudaMemcpyAsync(&n, e->d_scal, 4, cudaMemcpyDeviceToHost, S(stream)));
    AM_CUDA(cudaStreamSynchronize(S(stream)));
    out->min_x = e->u_min_x; out->max_x = e->u_max_x; out->min_y = e->u_min_y; out->max_y = e->u_max_y; out->size = e->u_size;
    out->crop_off = e->u_crop_off; out->arena = e->arena; out->n = n;
    return AM_OK;
}

extern "C" int am_est_export_sizes(am_estimator* e, long long* h_sizes, void* stream) {
    if (!e || !h_sizes) return AM_ERR_ARG;
    k_export_sizes<<<1, 1024, 0, S(stream)>>>(e->act[e->cur], e->d_scal, e->u_min_x, e->u_max_x, e->u_min_y, e->u_max_y, e->d_scal64 + 2);
    AM_CUDA(cudaMemcpyAsync(e->h_scal64 + 2, e->d_scal64 + 2, 16, cudaMemcpyDeviceToHost, S(stream)));
    AM_CUDA(cudaStreamSynchronize(S(stream)));
    h_sizes[0] = (long long)e->h_scal64[2]; h_sizes[1] = (long long)e->h_scal64[3];
    return AM_OK;
}
extern "C" int am_est_export(am_estimator* e, int* d_meta, uint32_t* d_crops, void* stream) {
    if (!e || !d_meta || !d_crops) return AM_ERR_ARG;
    k_export<<<1, 1024, 0, S(stream)>>>(e->act[e->cur], e->d_scal, e->u_min_x, e->u_max_x, e->u_min_y, e->u_max_y, e->u_size, e->u_last,
                                        e->u_first_frame, e->u_first_label, e->u_crop_off, e->arena, d_meta, d_crops, est_tmp(e));
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}
extern "C" int am_est_import(am_estimator* e, int n_active, int n_unique, int img_idx, unsigned long long tempo_count,
                             const int* d_meta, const uint32_t* d_crops, long long crop_words, void* stream) {
    if (!e || n_active < 0 || n_active > e->MA || n_unique > e->MU || (unsigned long long)crop_words > e->AW) return AM_ERR_CAPACITY;
    if (n_active > 0 && (!d_meta || !d_crops)) return AM_ERR_ARG;
    if (n_active > 0)
        k_import<<<1, 1024, 0, S(stream)>>>(n_active, d_meta, d_crops, e->act[e->cur], e->u_min_x, e->u_max_x, e->u_min_y, e->u_max_y,
                                            e->u_size, e->u_last, e->u_first_frame, e->u_first_label, e->u_crop_off, e->arena, est_tmp(e), e->MU);
    int sc[4] = {n_unique, n_active, img_idx, 0};
    unsigned long long sc64[2] = {tempo_count, (unsigned long long)crop_words};
    AM_CUDA(cudaMemcpyAsync(e->d_scal, sc, 16, cudaMemcpyHostToDevice, S(stream)));
    AM_CUDA(cudaMemcpyAsync(e->d_scal64, sc64, 16, cudaMemcpyHostToDevice, S(stream)));
    k_rebuild_boxes<<<64, 256, 0, S(stream)>>>(e->act[e->cur], e->d_scal, e->u_min_x, e->u_max_x, e->u_min_y, e->u_max_y, e->box[e->cur]);
    AM_CUDA(cudaStreamSynchronize(S(stream)));
    return AM_OK;
}

// Asynchronous hand-off: everything (sizes included) stays on the device, nothing synchronises the host.
extern "C" int am_est_export_dev(am_estimator* e, int* d_buf, long long capacity_words, void* stream) {
    if (!e || !d_buf || capacity_words < XHDR) return AM_ERR_ARG;
    k_export_dev_meta<<<1, 1024, 0, S(stream)>>>(e->act[e->cur], e->d_scal, e->d_scal64, e->u_min_x, e->u_max_x, e->u_min_y, e->u_max_y,
                                                 e->u_size, e->u_last, e->u_first_frame, e->u_first_label, d_buf, capacity_words, est_tmp(e));
    k_export_dev_crops<<<148, 256, 0, S(stream)>>>(e->act[e->cur], e->d_scal, e->u_min_x, e->u_max_x, e->u_min_y, e->u_max_y, e->u_crop_off,
                                                   e->arena, d_buf, capacity_words, est_tmp(e));
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}
extern "C" int am_est_import_dev(am_estimator* e, const int* d_buf, void* stream) {
    if (!e || !d_buf) return AM_ERR_ARG;
    k_import_dev_meta<<<1, 1024, 0, S(stream)>>>(d_buf, e->act[e->cur], e->d_scal, e->d_scal64, e->u_min_x, e->u_max_x, e->u_min_y, e->u_max_y,
                                                 e->u_size, e->u_last, e->u_first_frame, e->u_first_label, e->u_crop_off, e->MU, e->MA, e->AW);
    k_import_dev_crops<<<148, 256, 0, S(stream)>>>(d_buf, e->arena, e->AW);
    k_rebuild_boxes<<<64, 256, 0, S(stream)>>>(e->act[e->cur], e->d_scal, e->u_min_x, e->u_max_x, e->u_min_y, e->u_max_y, e->box[e->cur]);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}
