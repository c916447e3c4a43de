// k x k median filter (k = 3, 5), zero padding, over N = B*C planes.
//
// Replaces MiddleBlur.forward (noise_layers/middle_filter.py:11-13 -> kornia MedianBlur), which
// materialises a one-hot conv2d expansion of k*k times the image (288 / 672 B/px forward) and
// then sorts along it.  Here persistent CTAs stage halo tiles in a shared-memory ring (TMA; cp.async when the rows
// are not 16-byte aligned); a thread walks down a column strip, sorts each k-wide window ROW once (shared by the k
// vertically adjacent outputs) and selects the median from the k sorted rows with a pruned min/max network
// (3x3: FMNMX3 + middle-of-three, 5x5: generated median_pair_net.cuh) — pure selection, so the
// forward is bit-exact.  Optionally it records, per output, the raster position of the FIRST
// window element equal to the median (uint8, row stride idx_sh); the backward is a deterministic gather through it.
#include "median_pair_net.cuh"
#include "tma.cuh"
#include "wm_common.cuh"

namespace wm {

__device__ __forceinline__ float mid3(float a, float b, float c, float lo, float hi) {
    // the element that is neither the min nor the max: XOR of the five bit patterns
    return __int_as_float(__float_as_int(a) ^ __float_as_int(b) ^ __float_as_int(c) ^
                          __float_as_int(lo) ^ __float_as_int(hi));
}
__device__ __forceinline__ void sort3(float& a, float& b, float& c) {
    const float lo = fmin3(a, b, c), hi = fmax3(a, b, c);
    b = mid3(a, b, c, lo, hi); a = lo; c = hi;
}
__device__ __forceinline__ float med3(float a, float b, float c) {
    return mid3(a, b, c, fmin3(a, b, c), fmax3(a, b, c));
}
// (a == b) as 1.0f / 0.0f (FSET.BF): lets the arg-median search accumulate match bits with FFMA on
// the FMA pipe instead of FSETP + SEL pairs on the (half-rate, already saturated) ALU pipe
__device__ __forceinline__ float feq(float a, float b) {
    float d; asm("set.eq.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d;
}
// packed fp32x2 FMA (FFMA2): two independent IEEE fmas per issue slot
// a0 += g0 if the low halves of (h, w) are equal, a1 += g1 if the high halves are
__device__ __forceinline__ void add_if_eq2(float& a0, float& a1, uint32_t h, uint32_t w, float g0, float g1) {
    asm("{\n\t.reg .pred p, q;\n\tsetp.eq.f16x2 p|q, %2, %3;\n\t@p add.f32 %0, %0, %4;\n\t@q add.f32 %1, %1, %5;\n\t}"
        : "+f"(a0), "+f"(a1) : "r"(h), "r"(w), "f"(g0), "f"(g1));
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1,%2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 r;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rd));
    return r;
}
// code = sum_j match_j * 2^(n-1-j) (exact in fp32 for n <= 24): raster index of the FIRST match
__device__ __forceinline__ int first_match(float code, int n) {
    return n - 1 - ((__float_as_int(code) >> 23) - 127);
}

// max(a, b) as a + b - min(a, b) on the bit patterns: two IMADs (FMA pipe) instead of one FMNMX (ALU pipe); see the
// 5x5 kernel below.  one = +1 and neg1 = -1 must reach the kernel as launch parameters.
struct CeIntSum {
    int one, neg1;
    __device__ __forceinline__ void operator()(float& a, float& b) const {
        const float lo = fminf(a, b);
        int s, h;
        asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(s) : "r"(__float_as_int(a)), "r"(one), "r"(__float_as_int(b)));
        asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(h) : "r"(__float_as_int(lo)), "r"(neg1), "r"(s));
        b = __int_as_float(h); a = lo;
    }
};
template <class CE> __device__ __forceinline__ void sort5(float (&v)[5], const CE ce) {
    // optimal 9-comparator network
    ce(v[0], v[1]); ce(v[3], v[4]); ce(v[2], v[4]); ce(v[2], v[3]); ce(v[0], v[3]);
    ce(v[0], v[2]); ce(v[1], v[4]); ce(v[1], v[3]); ce(v[1], v[2]);
}
// the same network with its first NPLAIN comparators as plain min/max pairs (2 ALU-pipe instructions) and the rest as
// min + integer-sum max (1 ALU + 2 FMA-pipe instructions): a per-comparator balance between issue slots and the ALU pipe
#ifndef WM_M5_SORT_PLAIN
#define WM_M5_SORT_PLAIN 4
#endif
template <int NPLAIN, class CE> __device__ __forceinline__ void sort5_mixed(float (&v)[5], const CE ce) {
    const CeMinMax pl{};
    constexpr int A[9] = {0, 3, 2, 2, 0, 0, 1, 1, 1}, B[9] = {1, 4, 4, 3, 3, 2, 4, 3, 2};
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        if (i < NPLAIN) pl(v[A[i]], v[B[i]]); else ce(v[A[i]], v[B[i]]);
    }
}
template <bool INT> struct ce_pick { using type = CeMinMax; static __device__ __forceinline__ type make(int, int) { return type(); } };
template <> struct ce_pick<true> { using type = CeIntSum; static __device__ __forceinline__ type make(int one, int neg1) { return type{one, neg1}; } };


// ---------------------------------------------------------------------------------------------
// 3x3 fast path (16-byte aligned rows): persistent CTAs fed by a ring of TMA-staged halo tiles
// (out-of-bounds elements arrive as zeros == kornia's zero padding).  One warp per 8-row strip,
// one lane per 4 adjacent columns: per tile row a lane reads 6 values (LDS.128 + 2 LDS.32), sorts
// the 4 horizontal triples (FMNMX3 + XOR mid) and keeps the last 3 rows of sorted triples and raw
// values in registers; an output is max3(mins), med3(mids), min3(maxes) -> med3.  The arg-median
// plane is found by comparing the 9 raw values with the median (first match in raster order) and
// leaves as one 32-bit store per lane.
// ---------------------------------------------------------------------------------------------
constexpr int MT_TW = 128, MT_TH = 64, MT_HALO = 4, MT_BW = MT_TW + 2 * MT_HALO, MT_BH = MT_TH + 2,
              MT_THREADS = 256, MT_ROWS = 8, MT_STAGES = 3, MT_STRIDE = ((MT_BW * MT_BH + 31) / 32) * 32;

struct m_true { static constexpr bool value = true; };
struct m_false { static constexpr bool value = false; };

struct MedTArgs {
    float* y; uint8_t* idx; int N, H, W, tiles_x, tiles_y; int64_t total;
    StoreEp ep;
    int one, neg1;      // +1 / -1 as launch parameters (FMA-pipe integer sums, see CeIntSum)
    int64_t idx_sh;     // row stride of the arg-median plane in bytes (>= W; a multiple of 16 keeps the backward's idx ring on TMA)
    RaggedSrc rag;      // RAGGED instantiations: source planes (rows not 16-byte aligned, no tensor map)
};
// middle of three as a + b + c - min - max on the bit patterns: four IMADs on the FMA pipe instead of two LOP3 on the
// ALU pipe, which the arg-median search saturates (102 -> 96 us at 64x3x512x512 with the plane; without it the
// kernel is HBM-bound and keeps the XOR form)
__device__ __forceinline__ float mid3_sum(float a, float b, float c, float lo, float hi, int one, int neg1) {
    int s;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(s) : "r"(__float_as_int(a)), "r"(one), "r"(__float_as_int(b)));
    asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(s) : "r"(__float_as_int(c)), "r"(one));
    asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(s) : "r"(__float_as_int(lo)), "r"(neg1));
    asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(s) : "r"(__float_as_int(hi)), "r"(neg1));
    return __int_as_float(s);
}

// stage stride in elements: whole 128-byte lines for 4-byte and for 2-byte elements
// a TMA box must START on a 16-byte boundary of its row (a 2-byte box at an 8-byte offset is an illegal instruction),
// so the halo of a 2-byte tile is 8 elements
template <int DT> constexpr int mt_halo() { return DT == WM_DT_F32 ? MT_HALO : 8; }
template <int DT> constexpr int mt_bw() { return MT_TW + 2 * mt_halo<DT>(); }
template <int DT> constexpr int mt_stride() { return DT == WM_DT_F32 ? MT_STRIDE : ((mt_bw<DT>() * MT_BH + 63) / 64) * 64; }

// RAGGED (rows not 16-byte aligned): ring fed by cp.async, scalar stores (tma.cuh: stage_box_cpasync).
// IDT: element type of the source planes (float16 / bfloat16 tiles are widened on the way to registers: exact, so the
// median and its position are those of the float32 image); typed instantiations are TMA-only, no store epilogue.
template <bool WANT_IDX, bool EP, bool RAGGED = false, int IDT = WM_DT_F32>
__global__ void __launch_bounds__(MT_THREADS, 2) median3_tma_kernel(const __grid_constant__ CUtensorMap tmap, const MedTArgs a) {
    static_assert(IDT == WM_DT_F32 || (!RAGGED && !EP), "typed planes: TMA rows, plain store");
    constexpr int ES = tile_elem_size<IDT>(), SSTRIDE = mt_stride<IDT>(), HALO = mt_halo<IDT>(), BW = mt_bw<IDT>();
    extern __shared__ __align__(128) unsigned char tiles[];
    __shared__ uint64_t full[MT_STAGES];
    const int tid = threadIdx.x;
    if (!RAGGED && tid == 0) {
        tma_prefetch_desc(&tmap);
#pragma unroll
        for (int s = 0; s < MT_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int per_plane = a.tiles_x * a.tiles_y;
    auto issue = [&](int64_t t, int s) {
        const int n = int(t / per_plane), rem = int(t - int64_t(n) * per_plane);
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        if (RAGGED) {       // every thread; an empty group keeps the per-thread group count in step with the ring
            if (t < a.total) stage_box_cpasync<MT_THREADS>(reinterpret_cast<float*>(tiles) + s * MT_STRIDE, a.rag, n, a.H, a.W, tx * MT_TW - HALO, ty * MT_TH - 1, BW, MT_BH);
            else asm volatile("cp.async.commit_group;" ::: "memory");
        } else {
            mbar_expect_tx(&full[s], BW * MT_BH * ES);
            tma_load_3d(tiles + size_t(s) * SSTRIDE * ES, &tmap, tx * MT_TW - HALO, ty * MT_TH - 1, n, &full[s]);
        }
    };
    if (RAGGED || tid == 0) {
#pragma unroll
        for (int s = 0; s < MT_STAGES; ++s) {
            const int64_t t = int64_t(blockIdx.x) + int64_t(s) * gridDim.x;
            if (RAGGED || t < a.total) issue(t, s);
        }
    }
    const int cg = tid & 31, strip = tid >> 5;
    int it = 0;
    for (int64_t t = blockIdx.x; t < a.total; t += gridDim.x, ++it) {
        const int s = it % MT_STAGES;
        if (RAGGED) { cpasync_wait<MT_STAGES - 1>(); __syncthreads(); }
        else mbar_wait(&full[s], (it / MT_STAGES) & 1);
        const int n = int(t / per_plane), rem = int(t - int64_t(n) * per_plane);
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        const int gx = tx * MT_TW + 4 * cg, gy0 = ty * MT_TH + strip * MT_ROWS;
        const unsigned char* stage = tiles + size_t(s) * SSTRIDE * ES;
        const int col = (strip * MT_ROWS) * BW + HALO + 4 * cg;           // element index of the lane's first column
        float raw[3][6], lo[3][4], mi[3][4], hi[3][4];
        auto load_row = [&](int row, int slot) {
            const int p = col + row * BW;
            const float4 c = tile_ld4<IDT>(stage, p);
            raw[slot][0] = tile_ld1<IDT>(stage, p - 1); raw[slot][1] = c.x; raw[slot][2] = c.y; raw[slot][3] = c.z; raw[slot][4] = c.w;
            raw[slot][5] = tile_ld1<IDT>(stage, p + 4);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float u = raw[slot][c4], v = raw[slot][c4 + 1], w = raw[slot][c4 + 2];
                const float l = fmin3(u, v, w), h = fmax3(u, v, w);
                lo[slot][c4] = l; hi[slot][c4] = h; mi[slot][c4] = mid3(u, v, w, l, h);
            }
        };
        load_row(0, 0);
        load_row(1, 1);
        const bool col_ok = gx < a.W;
        const int64_t obase = (int64_t(n) * a.H + gy0) * a.W + gx;
        const int64_t ibase = (int64_t(n) * a.H + gy0) * a.idx_sh + gx;
        // FULL: this lane stores all MT_ROWS rows of its strip - no per-row test around the stores
        auto rows = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
        for (int r = 0; r < MT_ROWS; ++r) {
            load_row(r + 2, (r + 2) % 3);
            float4 o;
            float* op = &o.x;
            uint32_t packed = 0;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float m0 = mi[0][c4], m1 = mi[1][c4], m2 = mi[2][c4];
                const float A = fmax3(lo[0][c4], lo[1][c4], lo[2][c4]), C = fmin3(hi[0][c4], hi[1][c4], hi[2][c4]);
                float med;
                if (WANT_IDX) {
                    const float B = mid3_sum(m0, m1, m2, fmin3(m0, m1, m2), fmax3(m0, m1, m2), a.one, a.neg1);
                    med = mid3_sum(A, B, C, fmin3(A, B, C), fmax3(A, B, C), a.one, a.neg1);
                } else {
                    med = med3(A, med3(m0, m1, m2), C);
                }
                op[c4] = med;
                if (WANT_IDX) {
                    float code = 0.f;
#pragma unroll
                    for (int j = 0; j < 9; ++j)       // window row j/3 is ring slot (r + j/3) % 3
                        code = fmaf(code, 2.f, feq(raw[(r + j / 3) % 3][c4 + j % 3], med));
                    packed |= uint32_t(first_match(code, 9)) << (8 * c4);
                }
            }
            if (FULL || (col_ok && gy0 + r < a.H)) {
                if (EP) {   // the window's centre row is ring slot (r + 1) % 3: x is still in registers
                    const float* c = raw[(r + 1) % 3];
                    o = a.ep.from_input ? ep_apply4v(o, make_float4(c[1], c[2], c[3], c[4]), a.ep)
                                        : ep_apply4(o, a.ep.x + obase + int64_t(r) * a.W, a.ep);
                }
                if (RAGGED) st4_ragged(a.y + obase + int64_t(r) * a.W, o, gx, a.W);
                else stg128(a.y + obase + int64_t(r) * a.W, o);
                if (WANT_IDX) {
                    uint8_t* ip = a.idx + ibase + int64_t(r) * a.idx_sh;
                    if (!RAGGED || (a.idx_sh & 3) == 0) *reinterpret_cast<uint32_t*>(ip) = packed;   // bytes past W land in the row's padding
                    else {
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) if (gx + c4 < a.W) ip[c4] = uint8_t(packed >> (8 * c4));
                    }
                }
            }
        }
        };
        // (with the arg-median plane the kernel is issue-bound: 99.6 -> 95.7 us; without it it is HBM-bound and the second
        // copy of the loop only costs instruction cache: 73.8 -> 75.8 us, so that instantiation keeps one path)
        if (WANT_IDX && !RAGGED && col_ok && gy0 + MT_ROWS <= a.H) rows(m_true{}); else rows(m_false{});
        __syncthreads();
        if (RAGGED || tid == 0) {
            const int64_t t2 = t + int64_t(MT_STAGES) * gridDim.x;
            if (RAGGED || t2 < a.total) issue(t2, s);
        }
    }
}

// 5x5 fast path: same TMA ring, but a tile is 128 columns x 36 rows of TWO planes (one 3-D box): thread
// (c, half) owns column c of plane 2m + half (the older layout, two 36-row strips of one 72-row tile, computed
// ceil(H / 72) * 72 rows: 12.5 % more than the image at H = 512; this one at most 35 rows more, see m5_row0).
// A lane keeps the last 6 window rows sorted (9-comparator network per row, shared by the 5 vertically adjacent
// outputs) and produces TWO outputs per step: rows r and r+1 share four of their five window rows,
// whose 20 values are reduced once to the six of rank 7..12 — the only ones that can be the median of
// either window — and each output is then the median of those six and its own sorted row
// (median_pair_net.cuh, generated; the merged row pair (r+3, r+4) of one step is reused as the pair
// (r+1, r+2) of the next: 30 + 36 + 2 x 10 min/max per pair of outputs against 2 x 110 for the
// one-output-at-a-time network).  The raw rows stay in registers for the arg-median search.  The row loop
// is unrolled by 3 pairs so that ring slots are compile-time indices.
//
// The kernel is bound by the ALU pipe (FMNMX, FSET) and then by instruction issue, not by HBM (ncu: ALU 89 %, FMA
// 16 %, DRAM 14 %).  A compare-exchange therefore computes only its MIN with FMNMX; the max is a + b - min on the bit
// patterns, two IMADs on the otherwise idle FMA pipe (CeIntSum; the +1 / -1 multipliers are kernel parameters so
// that ptxas cannot fold them back into an ALU-pipe IADD3, and sit in uniform registers: two vector operands per
// IMAD, no register-bank stalls).  Bit-exact: the result is the other operand's bit pattern by construction.
constexpr int M5_TW = 128, M5_ROWS = 36, M5_HALO = 4, M5_BW = M5_TW + 2 * M5_HALO, M5_BH = M5_ROWS + 4,
              M5_THREADS = 256, M5_STAGES = 2, M5_STRIDE = 2 * M5_BW * M5_BH;
static_assert(M5_ROWS % 6 == 0 && (M5_STRIDE * sizeof(float)) % 128 == 0, "ring slots / stage alignment");

// First image row of tile row ty.  The row loop has a COMPILE-TIME trip count (with a data-dependent one ptxas moves the
// comparator multipliers out of the uniform registers and the kernel stalls on register banks: 294 -> 320 us), so the
// bottom tile is not cut short but shifted up to end at the image's last row; the rows it shares with the tile above
// are computed twice and written twice with identical bits.
__host__ __device__ __forceinline__ int m5_row0(int ty, int H) {
    const int r = ty * M5_ROWS, last = H > M5_ROWS ? H - M5_ROWS : 0;
    return r < last ? r : last;
}

template <int DT> constexpr int m5_halo() { return DT == WM_DT_F32 ? M5_HALO : 8; }
template <int DT> constexpr int m5_bw() { return M5_TW + 2 * m5_halo<DT>(); }
template <int DT> constexpr int m5_stride() { return 2 * m5_bw<DT>() * M5_BH; }
static_assert((m5_stride<WM_DT_F16>() * 2) % 128 == 0, "stage alignment of 2-byte tiles");

struct Med5Args {
    float* y; uint8_t* idx; int N, H, W, tiles_x, tiles_y; int64_t total;
    int one, neg1;
    StoreEp ep;
    int64_t idx_sh;     // row stride of the arg-median plane in bytes
    RaggedSrc rag;      // RAGGED instantiations: source planes
};

// One 128 x 36 tile of one plane: `col` = element index of the lane's window column in the staged box `stage`
// (row 0 = image row gy0 - 2; elements of type IDT).
// FULL: this lane stores all M5_ROWS rows of the tile (every lane inside the image when H >= M5_ROWS, thanks to the
// shifted bottom tile): the per-output row test and its branch go away.
template <bool WANT_IDX, bool EP, int IDT, bool FULL = false>
__device__ __forceinline__ void median5_tile(const Med5Args& a, const void* stage, int col, int64_t obase, int64_t ibase, int rows_ok) {
    // with the arg-median search on the ALU pipe too, every comparator moves its max to the FMA pipe; without it the
    // last network keeps plain FMNMX pairs (measured: 338 -> 261 us with, 231 -> 172 us without the plane at 64x3x504x512)
    constexpr int BW = m5_bw<IDT>();
    const auto ce_s = ce_pick<true>::make(a.one, a.neg1);
    const auto ce_c = ce_pick<WANT_IDX>::make(a.one, a.neg1);
    float srt[6][5], raw[WANT_IDX ? 6 : 1][5];
    auto load_row = [&](int row, int slot) {
        const int p = col + row * BW;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            srt[slot][k] = tile_ld1<IDT>(stage, p + k);
            if (WANT_IDX) raw[slot][k] = srt[slot][k];
        }
        if (WANT_IDX) sort5_mixed<WM_M5_SORT_PLAIN>(srt[slot], ce_s); else sort5(srt[slot], ce_s);
    };
#pragma unroll
    for (int j = 0; j < 4; ++j) load_row(j, j);
    float* const yb = a.y + obase;
    uint8_t* const ib = WANT_IDX ? a.idx + ibase : nullptr;
    int off = 0, ioff = 0;
    float mp[10];                                      // rows r+1, r+2 merged (kept from the previous step)
#pragma unroll
    for (int k = 0; k < 5; ++k) { mp[k] = srt[1][k]; mp[5 + k] = srt[2][k]; }
    merge10_sorted_5_5(mp, ce_s);
#pragma unroll 1
    for (int r0 = 0; r0 < M5_ROWS; r0 += 6) {
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            // outputs r and r + 1 share the window rows r+1 .. r+4 (ring slots are compile-time: r0 % 6 == 0)
            const int r = r0 + 2 * u;
            load_row(r + 4, (2 * u + 4) % 6);
            load_row(r + 5, (2 * u + 5) % 6);
            float mq[10], v[20];                       // rows r+3, r+4 merged: the next step's (r+1, r+2)
#pragma unroll
            for (int k = 0; k < 5; ++k) { mq[k] = srt[(2 * u + 3) % 6][k]; mq[5 + k] = srt[(2 * u + 4) % 6][k]; }
            merge10_sorted_5_5(mq, ce_s);
#pragma unroll
            for (int k = 0; k < 10; ++k) { v[k] = mp[k]; v[10 + k] = mq[k]; mp[k] = mq[k]; }
            mid6_of_2_sorted_10(v, ce_c);              // v[7..12]: the only shared values that can be a median
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                const int own = (2 * u + (o ? 5 : 0)) % 6;
                float w[11];
#pragma unroll
                for (int k = 0; k < 6; ++k) w[k] = v[7 + k];
#pragma unroll
                for (int k = 0; k < 5; ++k) w[6 + k] = srt[own][k];
                const float med = median11_sorted_6_5(w);
                int pos = 0;
                if (WANT_IDX) {
                    float hi = 0.f, lo = 0.f;          // rows 0-2 (15 bits) and rows 3-4 (10 bits)
#pragma unroll
                    for (int j = 0; j < 15; ++j)       // window row j/5 of output r + o is ring slot (2u + o + j/5) % 6
                        hi = fmaf(hi, 2.f, feq(raw[(2 * u + o + j / 5) % 6][j % 5], med));
#pragma unroll
                    for (int j = 15; j < 25; ++j)
                        lo = fmaf(lo, 2.f, feq(raw[(2 * u + o + j / 5) % 6][j % 5], med));
                    pos = hi != 0.f ? first_match(hi, 15) : 15 + first_match(lo, 10);
                }
                if (FULL || r + o < rows_ok) {
                    yb[off] = EP ? ep_apply(med, a.ep.from_input ? tile_ld1<IDT>(stage, col + (r + o + 2) * BW + 2) : a.ep.x[obase + off], a.ep) : med;
                    if (WANT_IDX) ib[ioff] = (uint8_t)pos;
                }
                off += a.W; ioff += int(a.idx_sh);
            }
        }
    }
}

template <bool WANT_IDX, bool EP, bool RAGGED = false, int IDT = WM_DT_F32>
__global__ void __launch_bounds__(M5_THREADS, 2) median5_tma_kernel(const __grid_constant__ CUtensorMap tmap, const Med5Args a) {
    static_assert(IDT == WM_DT_F32 || (!RAGGED && !EP), "typed planes: TMA rows, plain store");
    constexpr int ES = tile_elem_size<IDT>(), HALO = m5_halo<IDT>(), BW = m5_bw<IDT>(), STRIDE = m5_stride<IDT>();
    extern __shared__ __align__(128) unsigned char tiles[];
    __shared__ uint64_t full[M5_STAGES];
    const int tid = threadIdx.x;
    if (!RAGGED && tid == 0) {
        tma_prefetch_desc(&tmap);
#pragma unroll
        for (int s = 0; s < M5_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int per_plane = a.tiles_x * a.tiles_y;
    auto issue = [&](int64_t t, int s) {
        const int m = int(t / per_plane), rem = int(t - int64_t(m) * per_plane);
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        if (RAGGED) {       // both planes of the pair, one commit group (a plane past N arrives as zeros)
            if (t < a.total) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int n = 2 * m + hf;
                    float* dst = reinterpret_cast<float*>(tiles) + s * STRIDE + hf * (BW * M5_BH);
                    if (n < a.N) {
                        const float* plane = a.rag.x + int64_t(n) * a.rag.sp;
                        for (int i = threadIdx.x; i < BW * M5_BH; i += M5_THREADS) {
                            const int ly = i / BW, lx = i - ly * BW, gy = m5_row0(ty, a.H) - 2 + ly, gx = tx * M5_TW - HALO + lx;
                            const bool ok = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
                            const float* src = ok ? plane + int64_t(gy) * a.rag.sh + gx : plane;
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst + i)), "l"(src), "r"(ok ? 4 : 0) : "memory");
                        }
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        } else {
            mbar_expect_tx(&full[s], STRIDE * ES);
            tma_load_3d(tiles + size_t(s) * STRIDE * ES, &tmap, tx * M5_TW - HALO, m5_row0(ty, a.H) - 2, 2 * m, &full[s]);
        }
    };
    if (RAGGED || tid == 0) {
#pragma unroll
        for (int s = 0; s < M5_STAGES; ++s) {
            const int64_t t = int64_t(blockIdx.x) + int64_t(s) * gridDim.x;
            if (RAGGED || t < a.total) issue(t, s);
        }
    }
    const int c = tid & 127, half = tid >> 7;
    int it = 0;
    for (int64_t t = blockIdx.x; t < a.total; t += gridDim.x, ++it) {
        const int s = it % M5_STAGES;
        if (RAGGED) { cpasync_wait<M5_STAGES - 1>(); __syncthreads(); }
        else mbar_wait(&full[s], (it / M5_STAGES) & 1);
        const int m = int(t / per_plane), rem = int(t - int64_t(m) * per_plane);
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        const int n = 2 * m + half, gx = tx * M5_TW + c, gy0 = m5_row0(ty, a.H);
        const int rows_tile = min(M5_ROWS, a.H - gy0);
        const int col = half * (BW * M5_BH) + HALO - 2 + c;
        const int64_t obase = (int64_t(n) * a.H + gy0) * a.W + gx;
        const int rows_ok = (gx < a.W && n < a.N) ? rows_tile : 0;          // this lane stores output rows r < rows_ok
        // lanes that store the whole tile column (all of them inside an image of >= M5_ROWS rows) take the instantiation
        // without the per-output row test (-9 us at 64x3x512x512); lanes outside the image or the plane count do nothing
        if (rows_ok == M5_ROWS)
            median5_tile<WANT_IDX, EP, IDT, true>(a, tiles + size_t(s) * STRIDE * ES, col, obase, (int64_t(n) * a.H + gy0) * a.idx_sh + gx, rows_ok);
        else if (rows_ok > 0)
            median5_tile<WANT_IDX, EP, IDT>(a, tiles + size_t(s) * STRIDE * ES, col, obase, (int64_t(n) * a.H + gy0) * a.idx_sh + gx, rows_ok);
        __syncthreads();
        if (RAGGED || tid == 0) {
            const int64_t t2 = t + int64_t(M5_STAGES) * gridDim.x;
            if (RAGGED || t2 < a.total) issue(t2, s);
        }
    }
}

// Backward, TMA-fed (K = 3, 5): gx[p] = sum_{q in window(p)} [idx[q] == position of p in q's window] gy[q].
// Two tile rings (gy: float, idx: uint8 with a 16-byte halo); a lane owns 4 adjacent p columns and
// walks down its strip with the last K rows of (gy, idx) in registers.  Pure gather, fixed
// summation order: deterministic.
constexpr int MB_IBW = MT_TW + 32;
#ifndef WM_MB3_TH
#define WM_MB3_TH 64
#define WM_MB3_STAGES 2
#define WM_MB3_MINB 2
#endif
#ifndef WM_MB5_TH
#define WM_MB5_TH 64
#define WM_MB5_STAGES 2
#define WM_MB5_MINB 2
#endif
// tile rows, ring depth and CTAs per SM of the backward, per window size
template <int K> struct MBCfg {
    static constexpr int TH = K == 3 ? WM_MB3_TH : WM_MB5_TH, STAGES = K == 3 ? WM_MB3_STAGES : WM_MB5_STAGES,
                         MINB = K == 3 ? WM_MB3_MINB : WM_MB5_MINB;
    static constexpr int ROWS = TH / (MT_THREADS / 32);            // output rows per warp strip
    static constexpr int BH = TH + K - 1;
    static constexpr int GS = ((MT_BW * BH + 31) / 32) * 32;       // floats per cotangent stage
    static constexpr int IS = ((MB_IBW * BH + 127) / 128) * 128;   // bytes per idx stage
};

struct MedBArgs {
    void* gx; int N, H, W, tiles_x, tiles_y; int64_t total;
    RaggedSrc rag;      // RAGGED: the cotangent planes (rows not 16-byte aligned); the idx plane always has a tensor map
};

// RAGGED: the (gy) ring is filled by cp.async, the idx ring still by TMA (the forward wrote the arg-median plane with
// a 16-byte row stride), and gx leaves by scalar stores.
// ODT: element type of gx (float16 / bfloat16: the gradient leaves in the autocast type, rounded to nearest even).
template <int K, bool RAGGED = false, int ODT = WM_DT_F32>
__global__ void __launch_bounds__(MT_THREADS, MBCfg<K>::MINB) median_bwd_tma_kernel(const __grid_constant__ CUtensorMap tm_g,
                                                                       const __grid_constant__ CUtensorMap tm_i,
                                                                       const MedBArgs a) {
    using Cfg = MBCfg<K>;
    constexpr int R = K / 2, BH = Cfg::BH, GS = Cfg::GS, IS = Cfg::IS, WC = 4 + 2 * R, MB_STAGES = Cfg::STAGES, TH = Cfg::TH, ROWS = Cfg::ROWS;
    extern __shared__ __align__(128) float bufs[];
    uint8_t* ibufs = reinterpret_cast<uint8_t*>(bufs + MB_STAGES * GS);
    __shared__ uint64_t full[MB_STAGES];
    const int tid = threadIdx.x;
    if (tid == 0) {
        if (!RAGGED) tma_prefetch_desc(&tm_g);
        tma_prefetch_desc(&tm_i);
#pragma unroll
        for (int s = 0; s < MB_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int per_plane = a.tiles_x * a.tiles_y;
    auto issue = [&](int64_t t, int s) {
        const int n = int(t / per_plane), rem = int(t - int64_t(n) * per_plane);
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        if (RAGGED) {
            if (t < a.total) stage_box_cpasync<MT_THREADS>(bufs + s * GS, a.rag, n, a.H, a.W, tx * MT_TW - MT_HALO, ty * TH - R, MT_BW, BH);
            else asm volatile("cp.async.commit_group;" ::: "memory");
            if (tid == 0 && t < a.total) {
                mbar_expect_tx(&full[s], MB_IBW * BH);
                tma_load_3d(ibufs + s * IS, &tm_i, tx * MT_TW - 16, ty * TH - R, n, &full[s]);
            }
        } else {
            mbar_expect_tx(&full[s], MT_BW * BH * sizeof(float) + MB_IBW * BH);
            tma_load_3d(bufs + s * GS, &tm_g, tx * MT_TW - MT_HALO, ty * TH - R, n, &full[s]);
            tma_load_3d(ibufs + s * IS, &tm_i, tx * MT_TW - 16, ty * TH - R, n, &full[s]);
        }
    };
    if (RAGGED || tid == 0) {
#pragma unroll
        for (int s = 0; s < MB_STAGES; ++s) {
            const int64_t t = int64_t(blockIdx.x) + int64_t(s) * gridDim.x;
            if (RAGGED || t < a.total) issue(t, s);
        }
    }
    const int cg = tid & 31, strip = tid >> 5;
    int it = 0;
    for (int64_t t = blockIdx.x; t < a.total; t += gridDim.x, ++it) {
        const int s = it % MB_STAGES;
        if (RAGGED) { cpasync_wait<MB_STAGES - 1>(); __syncthreads(); }
        mbar_wait(&full[s], (it / MB_STAGES) & 1);
        const int n = int(t / per_plane), rem = int(t - int64_t(n) * per_plane);
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        const int gx = tx * MT_TW + 4 * cg, gy0 = ty * TH + strip * ROWS;
        const float* gcol = bufs + s * GS + (strip * ROWS) * MT_BW + MT_HALO + 4 * cg;
        const uint8_t* icol = ibufs + s * IS + (strip * ROWS) * MB_IBW + 16 + 4 * cg;
        float g[K][WC];                          // window columns -R .. 4+R-1 of the lane's 4 columns
        uint32_t hp[K == 5 ? K : 1][WC - 1];     // 5x5: (idx[j], idx[j+1]) of adjacent window columns as a half2 of 1024 + idx (0x64nn)
        int ix[K == 3 ? K : 1][WC];              // 3x3: the positions as integers
        auto load_row = [&](int row, int slot) {
            const float* p = gcol + row * MT_BW;
            const float4 c = *reinterpret_cast<const float4*>(p);
            g[slot][R] = c.x; g[slot][R + 1] = c.y; g[slot][R + 2] = c.z; g[slot][R + 3] = c.w;
            const uint8_t* q = icol + row * MB_IBW;
            if (K == 5) {
                const float2 l = *reinterpret_cast<const float2*>(p - 2), h = *reinterpret_cast<const float2*>(p + 4);
                g[slot][0] = l.x; g[slot][1] = l.y; g[slot][6] = h.x; g[slot][7] = h.y;
                const uint32_t* qw = reinterpret_cast<const uint32_t*>(q);
                const uint32_t wa = qw[-1], wb = qw[0], wc = qw[1], bias = 0x64646464u;
                const uint32_t u0 = __byte_perm(wa, wb, 0x5432), u1 = __byte_perm(wb, wc, 0x5432);   // window bytes 0..3, 4..7
                hp[slot][0] = __byte_perm(u0, bias, 0x4140); hp[slot][1] = __byte_perm(u0, bias, 0x4241);
                hp[slot][2] = __byte_perm(u0, bias, 0x4342); hp[slot][3] = __byte_perm(wb, bias, 0x4241);
                hp[slot][4] = __byte_perm(u1, bias, 0x4140); hp[slot][5] = __byte_perm(u1, bias, 0x4241);
                hp[slot][WC - 2] = __byte_perm(u1, bias, 0x4342);
            } else {
                g[slot][0] = p[-1]; g[slot][5] = p[4];
                const uint32_t w4 = *reinterpret_cast<const uint32_t*>(q);
                ix[slot][0] = q[-1]; ix[slot][1] = w4 & 0xff; ix[slot][2] = (w4 >> 8) & 0xff; ix[slot][3] = (w4 >> 16) & 0xff;
                ix[slot][4] = w4 >> 24; ix[slot][WC - 1] = q[4];
            }
        };
#pragma unroll
        for (int j = 0; j < K - 1; ++j) load_row(j, j);
        const bool col_ok = gx < a.W;
        const int64_t doff = (int64_t(n) * a.H + gy0) * a.W + gx;
        float* dst = reinterpret_cast<float*>(a.gx) + doff;
        // FULL: this lane stores all ROWS rows of its strip - no per-row test around the stores
        auto rows = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            load_row(r + K - 1, (r + K - 1) % K);
            float4 o;
            float* op = &o.x;
            // q = p + (dy, dx) is tile row r + R + dy = ring slot (r + R + dy) % K; p is at window position (R - dy, R - dx) of q
            if (K == 5) {
                // 25-term gather, two adjacent outputs per compare: one packed half compare (HSETP2) of the two positions
                // against the wanted one sets two predicates, each guarding a plain add - 1.5 instructions per term, fixed
                // summation order, nothing multiplied (a non-finite cotangent elsewhere in the window cannot leak in).
                // (Round 1's 2 FSET + 1 FFMA2 per pair of terms paid register-pair moves on top: 121 us against 94.)
#pragma unroll
                for (int c2 = 0; c2 < 4; c2 += 2) {
                    float a0 = 0.f, a1 = 0.f;
#pragma unroll
                    for (int dy = -R; dy <= R; ++dy)
#pragma unroll
                        for (int dx = -R; dx <= R; ++dx) {
                            const int slot = (r + R + dy) % K, col = c2 + R + dx;
                            const uint32_t want = 0x64006400u + 0x00010001u * uint32_t((R - dy) * K + (R - dx));
                            add_if_eq2(a0, a1, hp[slot][col], want, g[slot][col], g[slot][col + 1]);
                        }
                    op[c2] = a0; op[c2 + 1] = a1;
                }
            } else {
                // 9 terms: compare + select + add (the packed-compare form measured 3 us slower here: 85 against 82)
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    float acc = 0.f;
#pragma unroll
                    for (int dy = -R; dy <= R; ++dy)
#pragma unroll
                        for (int dx = -R; dx <= R; ++dx) {
                            const int slot = (r + R + dy) % K, want = (R - dy) * K + (R - dx);
                            acc += (ix[slot][c4 + R + dx] == want) ? g[slot][c4 + R + dx] : 0.f;
                        }
                    op[c4] = acc;
                }
            }
            if (FULL || (col_ok && gy0 + r < a.H)) {
                if (RAGGED) st4_ragged(dst + int64_t(r) * a.W, o, gx, a.W);
                else stg4_typed<ODT>(a.gx, doff + int64_t(r) * a.W, o);
            }
        }
        };
        // (5x5: issue-bound, 93.8 -> 90.5 us; the 3x3 backward is not, and keeps one path)
        if (K == 5 && !RAGGED && col_ok && gy0 + ROWS <= a.H) rows(m_true{}); else rows(m_false{});
        __syncthreads();
        if (RAGGED || tid == 0) {
            const int64_t t2 = t + int64_t(MB_STAGES) * gridDim.x;
            if (RAGGED || t2 < a.total) issue(t2, s);
        }
    }
}

template <int K, bool RAGGED, int ODT = WM_DT_F32>
static int launch_median_bwd_tma(const float* gy, const uint8_t* idx, int64_t idx_sh, void* gx, int N, int H, int W, cudaStream_t st) {
    CUtensorMap tg{}, ti;
    int rc = RAGGED ? 0 : tmap_planes(&tg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, gy, N, H, W, int64_t(H) * W, W, MT_BW, MBCfg<K>::BH);
    if (!rc) rc = tmap_planes(&ti, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, idx, N, H, W, int64_t(H) * idx_sh, idx_sh, MB_IBW, MBCfg<K>::BH);
    if (rc) { set_error("wm_median_bwd: cuTensorMapEncodeTiled failed (%d)", rc); return WM_E_ARG; }
    MedBArgs ba{gx, N, H, W, (W + MT_TW - 1) / MT_TW, (H + MBCfg<K>::TH - 1) / MBCfg<K>::TH, 0, RaggedSrc{gy, int64_t(H) * W, W}};
    ba.total = int64_t(N) * ba.tiles_x * ba.tiles_y;
    const size_t smem = size_t(MBCfg<K>::STAGES) * (sizeof(float) * MBCfg<K>::GS + MBCfg<K>::IS);
    cudaError_t e = cudaFuncSetAttribute(median_bwd_tma_kernel<K, RAGGED, ODT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "wm_median_bwd");
    const int64_t cap = int64_t(sm_count()) * MBCfg<K>::MINB;
    median_bwd_tma_kernel<K, RAGGED, ODT><<<(unsigned)(ba.total < cap ? ba.total : cap), MT_THREADS, smem, st>>>(tg, ti, ba);
    WM_LAUNCH_CHECK("wm_median_bwd(tma)");
    return WM_OK;
}

// gx[p] = sum over outputs q with p in window(q) and argmedian(q) == p of gy[q]
template <int K>
__global__ void __launch_bounds__(256) median_bwd_kernel(const float* __restrict__ gy, const uint8_t* __restrict__ idx,
                                                         int64_t idx_sh, float* __restrict__ gx, int N, int H, int W) {
    constexpr int R = K / 2;
    const int64_t total = int64_t(N) * H * W;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int w = int(i % W), h = int((i / W) % H);
        const int64_t base = i - (int64_t(h) * W + w);
        float acc = 0.f;
#pragma unroll
        for (int dy = -R; dy <= R; ++dy)
#pragma unroll
            for (int dx = -R; dx <= R; ++dx) {
                const int qh = h + dy, qw = w + dx;     // output whose window contains (h, w)
                if (qh >= 0 && qh < H && qw >= 0 && qw < W) {
                    // (h, w) sits at window position (R - dy, R - dx) of output (qh, qw)
                    const int want = (R - dy) * K + (R - dx);
                    const int64_t q = base + int64_t(qh) * W + qw;
                    if (idx[(base / W + qh) * idx_sh + qw] == want) acc += gy[q];
                }
            }
        gx[i] = acc;
    }
}

}  // namespace wm

using namespace wm;

template <int IDT>
static int launch_median3_typed(const CUtensorMap& tm, MedTArgs& ta, cudaStream_t st) {
    const size_t smem = size_t(tile_elem_size<IDT>()) * MT_STAGES * mt_stride<IDT>();
    auto kern = ta.idx ? median3_tma_kernel<true, false, false, IDT> : median3_tma_kernel<false, false, false, IDT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "wm_median_fwd_typed");
    const int64_t cap = int64_t(sm_count()) * 2;
    kern<<<(unsigned)(ta.total < cap ? ta.total : cap), MT_THREADS, smem, st>>>(tm, ta);
    WM_LAUNCH_CHECK("wm_median_fwd_typed(3x3)");
    return WM_OK;
}
template <int IDT>
static int launch_median5_typed(const CUtensorMap& tm, Med5Args& ta, cudaStream_t st) {
    const size_t smem = size_t(tile_elem_size<IDT>()) * M5_STAGES * m5_stride<IDT>();
    auto kern = ta.idx ? median5_tma_kernel<true, false, false, IDT> : median5_tma_kernel<false, false, false, IDT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "wm_median_fwd_typed");
    const int64_t cap = int64_t(sm_count()) * 2;
    kern<<<(unsigned)(ta.total < cap ? ta.total : cap), M5_THREADS, smem, st>>>(tm, ta);
    WM_LAUNCH_CHECK("wm_median_fwd_typed(5x5)");
    return WM_OK;
}

template <bool RAGGED>
static int launch_median3(const CUtensorMap& tm, MedTArgs& ta, cudaStream_t st) {
    const size_t smem = sizeof(float) * size_t(MT_STAGES) * MT_STRIDE;
    auto kern = ta.ep.x ? (ta.idx ? median3_tma_kernel<true, true, RAGGED> : median3_tma_kernel<false, true, RAGGED>)
                        : (ta.idx ? median3_tma_kernel<true, false, RAGGED> : median3_tma_kernel<false, false, RAGGED>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "wm_median_fwd");
    const int64_t cap = int64_t(sm_count()) * 2;
    kern<<<(unsigned)(ta.total < cap ? ta.total : cap), MT_THREADS, smem, st>>>(tm, ta);
    WM_LAUNCH_CHECK("wm_median_fwd(3x3)");
    return WM_OK;
}
template <bool RAGGED>
static int launch_median5(const CUtensorMap& tm, Med5Args& ta, cudaStream_t st) {
    const size_t smem = sizeof(float) * size_t(M5_STAGES) * M5_STRIDE;
    auto kern = ta.ep.x ? (ta.idx ? median5_tma_kernel<true, true, RAGGED> : median5_tma_kernel<false, true, RAGGED>)
                        : (ta.idx ? median5_tma_kernel<true, false, RAGGED> : median5_tma_kernel<false, false, RAGGED>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "wm_median_fwd");
    const int64_t cap = int64_t(sm_count()) * 2;
    kern<<<(unsigned)(ta.total < cap ? ta.total : cap), M5_THREADS, smem, st>>>(tm, ta);
    WM_LAUNCH_CHECK("wm_median_fwd(5x5)");
    return WM_OK;
}

extern "C" int wm_median_fwd(const float* x, int64_t x_sp, int64_t x_sh, float* y, uint8_t* idx, int64_t idx_sh,
                             int N, int H, int W, int k, const wm_store_epilogue* ep, void* stream) {
    if (N == 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(x && y, WM_E_NULL, "wm_median_fwd: null pointer");
    WM_EP_CHECK(ep, "wm_median_fwd");
    WM_REQUIRE(k == 3 || k == 5, WM_E_ARG, "wm_median_fwd: kernel size must be 3 or 5 (got %d)", k);
    WM_REQUIRE(N >= 0 && H > 0 && W > 0, WM_E_SHAPE, "wm_median_fwd: bad shape N=%d H=%d W=%d", N, H, W);
    WM_REQUIRE(!idx || idx_sh >= W, WM_E_ARG, "wm_median_fwd: the arg-median plane's row stride (%lld) must be >= W", (long long)idx_sh);
    cudaStream_t st = (cudaStream_t)stream;
    // rows on 16-byte boundaries: TMA-fed ring; anything else (W % 4 != 0, odd strides): the same kernels fed by cp.async
    const bool tma = W % 4 == 0 && aligned(y, 16) && tmap_ok(x, x_sp, x_sh, 4) && (k == 5 || !idx || (aligned(idx, 4) && idx_sh % 4 == 0));
    if (!tma) WM_EP_REJECT(ep, "wm_median_fwd (rows not 16-byte aligned)");
    CUtensorMap tm{};
    if (tma)
        if (int rc = tmap_planes(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, N, H, W, x_sp, x_sh, k == 3 ? MT_BW : M5_BW,
                                 k == 3 ? MT_BH : M5_BH, k == 3 ? 1 : 2)) {
            set_error("wm_median_fwd: cuTensorMapEncodeTiled failed (%d)", rc);
            return WM_E_ARG;
        }
    StoreEp sep = tma ? make_store_ep(ep) : StoreEp{nullptr, 0, 0, 0};
    sep.from_input = sep.x == x && x_sh == W && x_sp == int64_t(H) * W;
    if (k == 3) {
        MedTArgs ta{y, idx, N, H, W, (W + MT_TW - 1) / MT_TW, (H + MT_TH - 1) / MT_TH, 0, sep, 1, -1, idx_sh, RaggedSrc{x, x_sp, x_sh}};
        ta.total = int64_t(N) * ta.tiles_x * ta.tiles_y;
        return tma ? launch_median3<false>(tm, ta, st) : launch_median3<true>(tm, ta, st);
    }
    Med5Args ta{y, idx, N, H, W, (W + M5_TW - 1) / M5_TW, (H + M5_ROWS - 1) / M5_ROWS, 0, 1, -1, sep, idx_sh, RaggedSrc{x, x_sp, x_sh}};
    ta.total = int64_t((N + 1) / 2) * ta.tiles_x * ta.tiles_y;
    return tma ? launch_median5<false>(tm, ta, st) : launch_median5<true>(tm, ta, st);
}

extern "C" int wm_median_bwd(const float* gy, const uint8_t* idx, int64_t idx_sh, float* gx, int N, int H, int W, int k, void* stream) {
    if (N == 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(gy && idx && gx, WM_E_NULL, "wm_median_bwd: null pointer");
    WM_REQUIRE(k == 3 || k == 5, WM_E_ARG, "wm_median_bwd: kernel size must be 3 or 5 (got %d)", k);
    WM_REQUIRE(idx_sh >= W, WM_E_ARG, "wm_median_bwd: the arg-median plane's row stride (%lld) must be >= W", (long long)idx_sh);
    const int64_t total = int64_t(N) * H * W;
    if (total <= 0) return WM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (tmap_ok(idx, int64_t(H) * idx_sh, idx_sh, 1)) {         // idx rows on 16-byte boundaries (the forward's padded stride)
        if (W % 4 == 0 && aligned(gx, 16) && tmap_ok(gy, int64_t(H) * W, W, 4))
            return k == 3 ? launch_median_bwd_tma<3, false>(gy, idx, idx_sh, gx, N, H, W, st)
                          : launch_median_bwd_tma<5, false>(gy, idx, idx_sh, gx, N, H, W, st);
        return k == 3 ? launch_median_bwd_tma<3, true>(gy, idx, idx_sh, gx, N, H, W, st)
                      : launch_median_bwd_tma<5, true>(gy, idx, idx_sh, gx, N, H, W, st);
    }
    const int64_t want = (total + 255) / 256, cap = int64_t(sm_count()) * 32;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (k == 3) median_bwd_kernel<3><<<grid, 256, 0, st>>>(gy, idx, idx_sh, gx, N, H, W);
    else median_bwd_kernel<5><<<grid, 256, 0, st>>>(gy, idx, idx_sh, gx, N, H, W);
    WM_LAUNCH_CHECK("wm_median_bwd");
    return WM_OK;
}

// Typed planes (include/wm_attack.h): the forward reads float16 / bfloat16 source planes through the same TMA ring
// (exact widening: value and position are those of the float32 image), the backward stores gx in that type.
template <int IDT>
static int median_fwd_typed(const void* x, int64_t x_sp, int64_t x_sh, float* y, uint8_t* idx, int64_t idx_sh, int N, int H, int W, int k,
                            cudaStream_t st) {
    CUtensorMap tm{};
    if (int rc = tmap_planes(&tm, tile_tmap_type<IDT>(), tile_elem_size<IDT>(), x, N, H, W, x_sp, x_sh, k == 3 ? mt_bw<IDT>() : m5_bw<IDT>(),
                             k == 3 ? MT_BH : M5_BH, k == 3 ? 1 : 2)) {
        set_error("wm_median_fwd_typed: cuTensorMapEncodeTiled failed (%d)", rc);
        return WM_E_ARG;
    }
    const StoreEp sep{nullptr, 0, 0, 0};
    if (k == 3) {
        MedTArgs ta{y, idx, N, H, W, (W + MT_TW - 1) / MT_TW, (H + MT_TH - 1) / MT_TH, 0, sep, 1, -1, idx_sh, RaggedSrc{nullptr, 0, 0}};
        ta.total = int64_t(N) * ta.tiles_x * ta.tiles_y;
        return launch_median3_typed<IDT>(tm, ta, st);
    }
    Med5Args ta{y, idx, N, H, W, (W + M5_TW - 1) / M5_TW, (H + M5_ROWS - 1) / M5_ROWS, 0, 1, -1, sep, idx_sh, RaggedSrc{nullptr, 0, 0}};
    ta.total = int64_t((N + 1) / 2) * ta.tiles_x * ta.tiles_y;
    return launch_median5_typed<IDT>(tm, ta, st);
}

extern "C" int wm_median_fwd_typed(const void* x, int x_dtype, int64_t x_sp, int64_t x_sh, float* y, uint8_t* idx, int64_t idx_sh,
                                   int N, int H, int W, int k, void* stream) {
    if (N == 0) return WM_OK;
    if (x_dtype == WM_DT_F32) return wm_median_fwd(reinterpret_cast<const float*>(x), x_sp, x_sh, y, idx, idx_sh, N, H, W, k, nullptr, stream);
    WM_REQUIRE(x && y, WM_E_NULL, "wm_median_fwd_typed: null pointer");
    WM_REQUIRE(x_dtype == WM_DT_F16 || x_dtype == WM_DT_BF16, WM_E_ARG, "wm_median_fwd_typed: unknown element type %d", x_dtype);
    WM_REQUIRE(k == 3 || k == 5, WM_E_ARG, "wm_median_fwd_typed: kernel size must be 3 or 5 (got %d)", k);
    WM_REQUIRE(N > 0 && H > 0 && W > 0, WM_E_SHAPE, "wm_median_fwd_typed: bad shape N=%d H=%d W=%d", N, H, W);
    WM_REQUIRE(!idx || idx_sh >= W, WM_E_ARG, "wm_median_fwd_typed: the arg-median plane's row stride (%lld) must be >= W", (long long)idx_sh);
    WM_REQUIRE(W % 4 == 0 && aligned(y, 16) && tmap_ok(x, x_sp, x_sh, 2) && (k == 5 || !idx || (aligned(idx, 4) && idx_sh % 4 == 0)), WM_E_ALIGN,
               "wm_median_fwd_typed: 2-byte planes need rows on 16-byte boundaries (W %% 8 == 0, aligned strides); convert to float32 otherwise");
    cudaStream_t st = (cudaStream_t)stream;
    return x_dtype == WM_DT_F16 ? median_fwd_typed<WM_DT_F16>(x, x_sp, x_sh, y, idx, idx_sh, N, H, W, k, st)
                                : median_fwd_typed<WM_DT_BF16>(x, x_sp, x_sh, y, idx, idx_sh, N, H, W, k, st);
}

extern "C" int wm_median_bwd_typed(const float* gy, const uint8_t* idx, int64_t idx_sh, void* gx, int gx_dtype, int N, int H, int W, int k,
                                   void* stream) {
    if (N == 0) return WM_OK;
    if (gx_dtype == WM_DT_F32) return wm_median_bwd(gy, idx, idx_sh, reinterpret_cast<float*>(gx), N, H, W, k, stream);
    WM_REQUIRE(gy && idx && gx, WM_E_NULL, "wm_median_bwd_typed: null pointer");
    WM_REQUIRE(gx_dtype == WM_DT_F16 || gx_dtype == WM_DT_BF16, WM_E_ARG, "wm_median_bwd_typed: unknown element type %d", gx_dtype);
    WM_REQUIRE(k == 3 || k == 5, WM_E_ARG, "wm_median_bwd_typed: kernel size must be 3 or 5 (got %d)", k);
    WM_REQUIRE(N > 0 && H > 0 && W > 0 && idx_sh >= W, WM_E_SHAPE, "wm_median_bwd_typed: bad shape / idx stride");
    WM_REQUIRE(W % 4 == 0 && aligned(gx, 8) && tmap_ok(gy, int64_t(H) * W, W, 4) && tmap_ok(idx, int64_t(H) * idx_sh, idx_sh, 1), WM_E_ALIGN,
               "wm_median_bwd_typed: needs W %% 4 == 0, aligned planes and the forward's 16-byte idx rows; use wm_median_bwd + a cast otherwise");
    cudaStream_t st = (cudaStream_t)stream;
    if (gx_dtype == WM_DT_F16)
        return k == 3 ? launch_median_bwd_tma<3, false, WM_DT_F16>(gy, idx, idx_sh, gx, N, H, W, st)
                      : launch_median_bwd_tma<5, false, WM_DT_F16>(gy, idx, idx_sh, gx, N, H, W, st);
    return k == 3 ? launch_median_bwd_tma<3, false, WM_DT_BF16>(gy, idx, idx_sh, gx, N, H, W, st)
                  : launch_median_bwd_tma<5, false, WM_DT_BF16>(gy, idx, idx_sh, gx, N, H, W, st);
}
