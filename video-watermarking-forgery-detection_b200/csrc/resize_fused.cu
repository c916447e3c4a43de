// Fused Resize round trip:  y = clamp( up( down(x) ), 0, 1 )   and its exact adjoint, one kernel each.
//
// Replaces Resize.forward (noise_layers/resize.py:38-53): F.interpolate to int(r*H) x int(r*W),
// F.interpolate back to H x W, clamp — three full-tensor passes plus the materialised mid image
// (24 + 24 r^2 B/px) — by ONE kernel per direction that reads x once and writes y once (24 B/px;
// backward: gy + 1 bit/value mask in, gx out).
//
// Both interpolations are separable 4-tap gathers along each axis with F.interpolate's
// align_corners=False coordinates (ATen upsample_bicubic2d A=-0.75 / upsample_bilinear2d), taps
// index-clamped at the borders.  Along one axis the round trip out = U (D x) is therefore a BANDED
// n x n operator A = U D whose rows have at most floor(3 n/nm) + 5 non-zeros (7 for r = 1.5, 11
// for r = 0.5).  A small table kernel builds, per axis, the band start and weights of every row of
// A (forward) and of A^T (adjoint) with ATen's fp32 coordinate arithmetic; the main kernel is a
// generic separable banded transform  Y = A_v X A_h^T  on shared-memory tiles:
//   stage  : TMA box of the source region of a tile (backward: cotangent, then masked in place)
//   H pass : lane = output column (its band weights live in registers), warps walk the rows
//   V pass : lane = 4 adjacent columns (LDS.128), warps walk output row QUADS that share one window of tmp
//            (each tmp value is read (BT+2)/4 times instead of BT), broadcast weights,
//            clamp + 1-bit pass-through mask (4 ballots per tile row), STG.128
// CTAs are persistent over the planes of one tile position, so tables are read once per CTA.
// The adjoint is the SAME kernel run with the tables of A^T: a deterministic gather, no atomics
// (ATen's upsample backward uses atomicAdd).  A first version streamed each line through 4-wide
// register windows (one LDS per value); it was issue-bound on its data-dependent control flow
// (profiles/ncu_r1_resize_stream.txt).
#include "tma.cuh"
#include "wm_common.cuh"

namespace wm {

constexpr int RB_TW = 128, RB_TH = 64, RB_THREADS = 256;
#ifndef WM_RB_FULL_PATH
#define WM_RB_FULL_PATH 1
#endif
constexpr bool RB_FULL_PATH = WM_RB_FULL_PATH;
struct wm_true { static constexpr bool value = true; };
struct wm_false { static constexpr bool value = false; };
constexpr float RB_RATIO_MIN = 0.45f, RB_RATIO_MAX = 2.2f;

__device__ __forceinline__ float rb_cubic1(float x) { const float A = -0.75f; return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float rb_cubic2(float x) { const float A = -0.75f; return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

// Packed fp32x2 FMA (Blackwell FFMA2): two independent IEEE fmas per issue slot, bit-identical to fmaf.
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

// Taps of output index o: positions i0-1 .. i0+2 of the edge-replicated source, weights w.
// Same fp32 coordinate arithmetic as ATen (area_pixel_compute_source_index, align_corners=false).
template <int MODE>
__device__ __forceinline__ void rb_tap(float scale, int o, int n_in, int& i0, float (&w)[4]) {
    float rho = scale * (o + 0.5f) - 0.5f;
    if (MODE == 0) {
        rho = fmaxf(rho, 0.f);
        i0 = min(int(rho), n_in - 1);
        const float l1 = fminf(fmaxf(rho - i0, 0.f), 1.f);
        w[0] = 0.f; w[1] = 1.f - l1; w[2] = l1; w[3] = 0.f;
    } else {
        const float fl = floorf(rho);
        i0 = int(fl);
        const float t = rho - fl;
        w[0] = rb_cubic2(t + 1.f); w[1] = rb_cubic1(t); w[2] = rb_cubic1(1.f - t); w[3] = rb_cubic2(2.f - t);
    }
}

// ---------------------------------------------------------------------------------------------
// tables: per axis and direction  lo[n] (first source index of the band) + w[n][BT]
// workspace layout (floats): [fwd x | fwd y | adj x | adj y], each  n ints  +  n * BT floats
// ---------------------------------------------------------------------------------------------
struct RBAxis { int n, nm; float sd, su; };       // size, mid size, down scale n/nm, up scale nm/n

__host__ __device__ inline int64_t rb_axis_words(int n, int BT) { return int64_t(n) * (BT + 1); }

template <int MODE>
__global__ void rb_fwd_tables_kernel(RBAxis ax, int BT, int* __restrict__ lo, float* __restrict__ w) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= ax.n) return;
    int iu0; float wu[4];
    rb_tap<MODE>(ax.su, o, ax.nm, iu0, wu);
    int first = 0;
    for (int j = 0; j < BT; ++j) w[int64_t(o) * BT + j] = 0.f;
    for (int k = 0; k < 4; ++k) {
        const int mq = min(max(iu0 - 1 + k, 0), ax.nm - 1);
        int id0; float wd[4];
        rb_tap<MODE>(ax.sd, mq, ax.n, id0, wd);
        if (k == 0) first = min(max(id0 - 1, 0), ax.n - 1);      // indices are non-decreasing in (k, j)
        for (int j = 0; j < 4; ++j) {
            const int c = min(max(id0 - 1 + j, 0), ax.n - 1) - first;
            if (c >= 0 && c < BT) w[int64_t(o) * BT + c] += wu[k] * wd[j];
        }
    }
    lo[o] = first;
}

// adjoint rows: for input index i the outputs o with A[o][i] != 0 form a contiguous range
__device__ __forceinline__ int rb_first_output(int n, int BT, const int* __restrict__ flo, const float* __restrict__ fw, int i) {
    for (int o = max(0, i - 2 * BT - 8); o <= min(n - 1, i + 2 * BT + 8); ++o) {
        const int c = i - flo[o];
        if (c >= 0 && c < BT && fw[int64_t(o) * BT + c] != 0.f) return o;
    }
    return -1;
}

__global__ void rb_adj_tables_kernel(int n, int BT, const int* __restrict__ flo, const float* __restrict__ fw,
                                     int* __restrict__ lo, float* __restrict__ w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int j = 0; j < BT; ++j) w[int64_t(i) * BT + j] = 0.f;
    int first = rb_first_output(n, BT, flo, fw, i);
    if (first >= 0) {
        for (int o = first; o < min(first + BT, n); ++o) {
            const int c = i - flo[o];
            if (c >= 0 && c < BT) w[int64_t(i) * BT + (o - first)] = fw[int64_t(o) * BT + c];
        }
    } else {
        // a source index no output samples (bilinear downsampling skips some): all-zero row; keep the
        // band starts non-decreasing by inheriting the start of the nearest sampled index below
        for (int k = i - 1; k >= 0 && first < 0; --k) first = rb_first_output(n, BT, flo, fw, k);
        if (first < 0) first = 0;
    }
    lo[i] = first;
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
struct RBArgs {
    const void* src; int64_t s_sp, s_sh;          // source planes (backward: dense float32 cotangent); forward: element type DT
    void* dst;                                    // dense [N, H, W]; backward: element type DT
    uint32_t* mask;                               // [N, H, tiles_x, 4] ballot words (see header)
    const int* lox; const float* wx;              // column-axis tables of this direction
    const int* loy; const float* wy;              // row-axis tables
    int N, H, W, tiles_x, tiles_y;
    int* overflow;                                // optional: set to 1 if a band did not fit its window
    StoreEp ep;                                   // forward only: store epilogue (x dense, same layout as dst)
};

template <int BT> struct RBGeom {
    static constexpr int IW = ((RB_TW + BT + 20 + 3) / 4) * 4;      // staged columns (start = 4-aligned band start - 8)
    static constexpr int IH = RB_TH + BT + 4;                       // staged rows
    static constexpr int IHA = ((IH + 7) / 8) * 8;                  // allocated rows (the H pass runs 8 rows per step)
    static constexpr int NQ = RB_TH / 4;                            // output row quads of a tile
    static constexpr int BTV = BT + 2;                              // window shared by the 4 rows of a quad
    static constexpr size_t smem = sizeof(float) * (size_t(IHA) * IW + size_t(IHA) * RB_TW + size_t(NQ) * 4 * BTV) +
                                   sizeof(int) * NQ + sizeof(uint32_t) * IH * 12 + 128;
    // staged columns of a 2-byte source tile: a TMA box must start on a 16-byte boundary of its row (a 2-byte box at an
    // 8-byte offset is an illegal instruction), so its first column is rounded down to a multiple of 8 (up to 4 columns
    // further left than the float32 tile) and its rows are whole 16-byte groups: a superset of the float32 tile's
    // columns, so the proof of fit holds.  IW2 * 2 bytes <= IW * 4: the tile lives in the same region.
    static constexpr int IW2 = ((IW + 4 + 7) / 8) * 8;
};

// DIR 0: forward (clamp, mask out)   DIR 1: adjoint (cotangent masked in shared memory, no clamp)
// RAGGED: rows that are not 16-byte aligned (W % 4 != 0, odd strides) have no tensor map: the source tile is staged by
// 4-byte cp.async from every thread (zero fill outside the plane) and the output leaves by scalar stores; the mask
// words (rows of whole 16-byte groups for any W) stay on TMA.
// DT: float16 / bfloat16 at the boundary - the forward's SOURCE planes (staged as they are, widened in the H pass) or
// the adjoint's RESULT (stored rounded to nearest even); TMA rows only, no store epilogue.
template <int BT, int DIR, bool RAGGED = false, int DT = WM_DT_F32>
__global__ void __launch_bounds__(RB_THREADS, 2) rb_banded_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_mask,
                                                                  const RBArgs a) {
    using G = RBGeom<BT>;
    static_assert(DT == WM_DT_F32 || !RAGGED, "typed planes need 16-byte rows");
    constexpr int IDT = DIR == 0 ? DT : WM_DT_F32, ODT = DIR == 1 ? DT : WM_DT_F32;
    constexpr int IWS = IDT == WM_DT_F32 ? G::IW : G::IW2;           // row pitch of the staged source tile (elements)
    constexpr bool PLAIN = DT == WM_DT_F32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* in = reinterpret_cast<float*>(smem_raw);                  // [IHA][IW] (TMA fills IH rows; IDT elements, pitch IWS)
    float* tmp = in + G::IHA * G::IW;                                // [IHA][TW]
    uint32_t* mb = reinterpret_cast<uint32_t*>(tmp + G::IHA * RB_TW);   // [IH][12] mask words of the staged rows (adjoint)
    float* wyq = reinterpret_cast<float*>(mb + G::IH * 12);          // [NQ][BTV][4] weights of a row quad on its shared window
    int* yloq = reinterpret_cast<int*>(wyq + G::NQ * 4 * G::BTV);    // [NQ] first staged row of the quad's window
    __shared__ uint64_t full;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = blockIdx.x, ty = blockIdx.y;
    const int ox0 = tx * RB_TW, oy0 = ty * RB_TH;
    const int tw = min(RB_TW, a.W - ox0), th = min(RB_TH, a.H - oy0);

    // staged region: columns [xs, xs + IW) (xs may be negative: zero filled), rows [ys, ys + IH).
    // The 8-column margin absorbs the regularised window starts (flat band starts at a clamped border).
    const int xs = (__ldg(a.lox + ox0) & (IDT == WM_DT_F32 ? ~3 : ~7)) - 8;
    const int ys = __ldg(a.loy + oy0);
    // rows the V pass can touch (rows past the image bottom are staged as zeros: their weights are 0)
    const int ih = min(max(__ldg(a.loy + oy0 + th - 1) + BT, __ldg(a.loy + oy0 + ((th - 1) & ~3)) + BT + 2) - ys, G::IH);

    // H pass role: lane = TWO adjacent output columns whose bands share one window of BTW source
    // values, read as BTW/2 LDS.64.  The 32 windows of a warp are made REGULAR (start = even base +
    // 2*lane: the 8-byte accesses of a half-warp hit 32 distinct banks); each output's band is shifted
    // inside the window accordingly.
    constexpr int BTW = BT <= 10 ? 10 : 14;
    const int hg = warp & 1, hr = warp >> 1;                         // 64-column group, row set
    const int oa = 64 * hg + 2 * lane;                               // tile-local column of slot 0
    const bool hv0 = oa < tw, hv1 = oa + 1 < tw;
    const int hg0 = min(ox0 + oa, a.W - 1), hg1 = min(ox0 + oa + 1, a.W - 1);
    const int hlo0 = __ldg(a.lox + hg0), hlo1 = __ldg(a.lox + hg1);
    int hmin = hv0 ? hlo0 - 2 * lane : (1 << 30);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) hmin = min(hmin, __shfl_xor_sync(0xffffffffu, hmin, d));
    hmin = min(hmin, a.W) & ~1;                                      // even (also for negative values)
    const int hs = hmin + 2 * lane;
    const int hshift0 = hv0 ? hlo0 - hs : BTW, hshift1 = hv1 ? hlo1 - hs : BTW;
    float w0[BTW], w1[BTW];
#pragma unroll
    for (int t = 0; t < BTW; ++t) {
        const int j0 = t - hshift0, j1 = t - hshift1;
        w0[t] = (j0 >= 0 && j0 < BT) ? __ldg(a.wx + int64_t(hg0) * BT + j0) : 0.f;
        w1[t] = (j1 >= 0 && j1 < BT) ? __ldg(a.wx + int64_t(hg1) * BT + j1) : 0.f;
    }
    if (a.overflow) {                    // a weight that does not fit the shared window would be lost
        bool lost = false;
        for (int j = 0; j < BT; ++j) {
            lost |= hv0 && j + hshift0 >= BTW && __ldg(a.wx + int64_t(hg0) * BT + j) != 0.f;
            lost |= hv1 && j + hshift1 >= BTW && __ldg(a.wx + int64_t(hg1) * BT + j) != 0.f;
        }
        if (lost) atomicExch(a.overflow, 1);
    }
    const int hbase = min(max(hs - xs, 0), IWS - BTW) & ~1;
    if (a.overflow && hv0 && hbase != hs - xs) atomicExch(a.overflow, 1);

    // V pass tables: rows 4p .. 4p+3 share the BTV-row window that starts at the first row's band start
    constexpr int BTV = G::BTV;
    for (int i = tid; i < G::NQ * BTV * 4; i += RB_THREADS) {
        const int p = i / (BTV * 4), t = (i / 4) % BTV, k = i & 3;
        const int r0 = min(oy0 + 4 * p, a.H - 1), rk = min(oy0 + 4 * p + k, a.H - 1);
        const int d = __ldg(a.loy + rk) - __ldg(a.loy + r0);
        const int j = t - d;
        wyq[i] = (j >= 0 && j < BT) ? __ldg(a.wy + int64_t(rk) * BT + j) : 0.f;
        if (a.overflow && t < BT && t + d >= BTV && __ldg(a.wy + int64_t(rk) * BT + t) != 0.f) atomicExch(a.overflow, 1);
    }
    if (tid < G::NQ) yloq[tid] = min(max(__ldg(a.loy + min(oy0 + 4 * tid, a.H - 1)) - ys, 0), G::IH - BTV);
    if (tid == 0) {
        if (!RAGGED) tma_prefetch_desc(&tmap);
        mbar_init(&full, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const bool masked = DIR == 1 && a.mask != nullptr;
    const int mt0 = max(xs, 0) >> 7;                                 // first 128-column mask tile of the staged region
    auto request = [&](int plane) {        // TMA: thread 0 only.  RAGGED: every thread (its share of the cp.async copies)
        if (RAGGED) {
            stage_box_cpasync<RB_THREADS>(in, RaggedSrc{reinterpret_cast<const float*>(a.src), a.s_sp, a.s_sh}, plane, a.H, a.W, xs, ys, G::IW, G::IH);
            if (masked && tid == 0) {
                mbar_expect_tx(&full, 12 * G::IH * sizeof(uint32_t));
                tma_load_3d(mb, &tmap_mask, 4 * mt0, ys, plane, &full);
            }
        } else {
            mbar_expect_tx(&full, IWS * G::IH * tile_elem_size<IDT>() + (masked ? 12 * G::IH * sizeof(uint32_t) : 0));
            tma_load_3d(in, &tmap, xs, ys, plane, &full);
            if (masked) tma_load_3d(mb, &tmap_mask, 4 * mt0, ys, plane, &full);
        }
    };
    int n = blockIdx.z;
    if ((RAGGED || tid == 0) && n < a.N) request(n);
    for (int it = 0; n < a.N; n += gridDim.z, ++it) {
        // ---- stage: the box of this plane was requested one plane ago ---------------------------
        if (RAGGED) { cpasync_wait<0>(); __syncthreads(); }
        if (!RAGGED || masked) mbar_wait(&full, it & 1);
        if (masked) {
            // gy .* mask in place; the 4 ballot words of a (row, 128-column tile) sit in one uint4 of mb
            constexpr int C4 = G::IW / 4;
            for (int i = tid; i < ih * C4; i += RB_THREADS) {
                const int r = i / C4, c4 = i - r * C4;
                const int gx = xs + 4 * c4;
                if (gx >= 0 && gx < a.W) {
                    float4* q = reinterpret_cast<float4*>(in + r * G::IW + 4 * c4);
                    float4 v = *q;
                    const uint4 m = *reinterpret_cast<const uint4*>(mb + r * 12 + 4 * ((gx >> 7) - mt0));
                    const int b = (gx & 127) >> 2;
                    v.x = (m.x >> b) & 1u ? v.x : 0.f; v.y = (m.y >> b) & 1u ? v.y : 0.f;
                    v.z = (m.z >> b) & 1u ? v.z : 0.f; v.w = (m.w >> b) & 1u ? v.w : 0.f;
                    *q = v;
                }
            }
            __syncthreads();
        }

        // ---- H pass: tmp[r][oa .. oa+1] = sum_t {w0, w1}[t] * in[r][hbase + t] ----------------------
        {
            int p = hr * IWS + hbase;                      // element index in the staged tile
            float* q = tmp + hr * RB_TW + oa;
            for (int r = hr; r < ih; r += 8) {            // rows r and r + 4 (rows past ih land in spare rows)
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int pk = p + 4 * k * IWS;
                    float2 a0 = make_float2(0.f, 0.f), a1 = a0;        // (even taps, odd taps) partial sums
#pragma unroll
                    for (int t = 0; t < BTW; t += 2) {
                        const float2 v = tile_ld2<IDT>(in, pk + t);
                        a0 = ffma2(make_float2(w0[t], w0[t + 1]), v, a0);
                        a1 = ffma2(make_float2(w1[t], w1[t + 1]), v, a1);
                    }
                    *reinterpret_cast<float2*>(q + 4 * k * RB_TW) = make_float2(a0.x + a0.y, a1.x + a1.y);
                }
                p += 8 * IWS; q += 8 * RB_TW;
            }
        }
        __syncthreads();
        // `in` is dead: prefetch the next plane's tile while the V pass runs
        if ((RAGGED || tid == 0) && n + gridDim.z < a.N) request(n + gridDim.z);

        // ---- V pass: rows 4p .. 4p+3 x 4 columns per lane from one shared window of tmp --------------
        // FULL (a whole 128 x 64 tile inside the image): no column / row tests around the stores
        auto vpass = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            const bool okc = FULL || 4 * lane < tw;
            const bool want_mask = DIR == 0 && a.mask != nullptr;
            // running pointers of this warp's row quads (q = warp, warp + 8, ...)
            int64_t doff = (int64_t(n) * a.H + oy0 + 4 * warp) * a.W + ox0 + 4 * lane;     // element offset of this lane's first output
            float* const dst32 = reinterpret_cast<float*>(a.dst);
            uint32_t* mrow = a.mask + ((int64_t(n) * a.H + oy0 + 4 * warp) * a.tiles_x + tx) * 4 + (lane & 3);
            const int drow_step = 4 * (RB_THREADS / 32) * a.W, mrow_step = 4 * (RB_THREADS / 32) * a.tiles_x * 4;
            const int mrow_k = a.tiles_x * 4;
            for (int qd = warp; 4 * qd < th; qd += RB_THREADS / 32, doff += drow_step, mrow += mrow_step) {
                const float4* wq = reinterpret_cast<const float4*>(wyq + qd * BTV * 4);     // [t] -> weights of the 4 rows
                const float* p = tmp + yloq[qd] * RB_TW + 4 * lane;
                float2 lo[4], hi[4];                                // columns (0,1), (2,3) of the 4 rows
#pragma unroll
                for (int k = 0; k < 4; ++k) lo[k] = hi[k] = make_float2(0.f, 0.f);
                // store epilogue: x at the output positions is requested now (L2 hits: the tile was just
                // staged from the same lines) so that its latency hides under the window loop
                float4 xe[4];
                if (PLAIN && !RAGGED && DIR == 0 && a.ep.x) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        xe[k] = (okc && (FULL || 4 * qd + k < th)) ? ldg128_nc(a.ep.x + doff + k * a.W) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int t = 0; t < BTV; ++t) {
                    const float4 v = *reinterpret_cast<const float4*>(p + t * RB_TW);
                    const float4 w4 = wq[t];
                    const float2 vl = make_float2(v.x, v.y), vh = make_float2(v.z, v.w);
                    const float wk[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 wp = make_float2(wk[k], wk[k]);
                        lo[k] = ffma2(wp, vl, lo[k]); hi[k] = ffma2(wp, vh, hi[k]);
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool okr = FULL || 4 * qd + k < th;      // uniform
                    const float4 acc = make_float4(lo[k].x, lo[k].y, hi[k].x, hi[k].y);
                    if (DIR == 0) {
                        const float4 c = clamp01_nan4(acc);
                        if (RAGGED) { if (okc && okr) st4_ragged(dst32 + doff + k * a.W, c, ox0 + 4 * lane, a.W); }
                        else if (okc && okr) stg128(dst32 + doff + k * a.W, (PLAIN && a.ep.x) ? ep_apply4v(c, xe[k], a.ep) : c);
                        if (want_mask) {   // 0 <= v <= 1  <=>  saturate(v) == v  (false for NaN)
                            const unsigned b0 = __ballot_sync(0xffffffffu, okc && c.x == acc.x);
                            const unsigned b1 = __ballot_sync(0xffffffffu, okc && c.y == acc.y);
                            const unsigned b2 = __ballot_sync(0xffffffffu, okc && c.z == acc.z);
                            const unsigned b3 = __ballot_sync(0xffffffffu, okc && c.w == acc.w);
                            unsigned w = b0;                       // lane l < 4 stores word l
                            w = (lane & 3) == 1 ? b1 : w; w = (lane & 3) == 2 ? b2 : w; w = (lane & 3) == 3 ? b3 : w;
                            if (lane < 4 && okr) mrow[k * mrow_k] = w;
                        }
                    } else if (okc && okr) {
                        if (RAGGED) st4_ragged(dst32 + doff + k * a.W, acc, ox0 + 4 * lane, a.W);
                        else stg4_typed<ODT>(a.dst, doff + k * a.W, acc);
                    }
                }
            }
        };
        // (forward, TMA rows: -2..3 us per launch at 64x3x512x512; the adjoint's plain stores gain nothing)
        if (RB_FULL_PATH && DIR == 0 && !RAGGED && tw == RB_TW && th == RB_TH) vpass(wm_true{}); else vpass(wm_false{});
        __syncthreads();          // tmp is rewritten by the next plane's H pass
    }
}

// Proof that a geometry fits: the SAME window arithmetic as rb_banded_kernel's prologue, evaluated from the tables
// alone for every tile position (one CTA each), for one direction's table set.  Sets *flag when a non-zero band
// weight would fall outside the lane pair's register window (H pass), the staged columns, the row quad's shared
// window (V pass) or the staged rows.  Launched by wm_resize_tables for both directions, so the flag in the
// workspace's last word is final when the tables are: the host reads it ONCE per new geometry.
template <int BT>
__global__ void __launch_bounds__(RB_THREADS) rb_prove_kernel(const int* __restrict__ lox, const float* __restrict__ wx,
                                                              const int* __restrict__ loy, const float* __restrict__ wy,
                                                              int H, int W, int* __restrict__ flag) {
    using G = RBGeom<BT>;
    constexpr int BTW = BT <= 10 ? 10 : 14, BTV = G::BTV;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ox0 = blockIdx.x * RB_TW, oy0 = blockIdx.y * RB_TH;
    const int tw = min(RB_TW, W - ox0), th = min(RB_TH, H - oy0);
    const int xs = (lox[ox0] & ~3) - 8, ys = loy[oy0];
    bool bad = false;
    if (warp < 2) {                                   // the two 64-column groups of the H pass
        const int oa = 64 * warp + 2 * lane;
        const bool hv0 = oa < tw, hv1 = oa + 1 < tw;
        const int hg0 = min(ox0 + oa, W - 1), hg1 = min(ox0 + oa + 1, W - 1);
        const int hlo0 = lox[hg0], hlo1 = lox[hg1];
        int hmin = hv0 ? hlo0 - 2 * lane : (1 << 30);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) hmin = min(hmin, __shfl_xor_sync(0xffffffffu, hmin, d));
        hmin = min(hmin, W) & ~1;
        const int hs = hmin + 2 * lane;
        const int hshift0 = hv0 ? hlo0 - hs : BTW, hshift1 = hv1 ? hlo1 - hs : BTW;
        for (int j = 0; j < BT; ++j) {
            bad |= hv0 && (j + hshift0 >= BTW || j + hshift0 < 0) && wx[int64_t(hg0) * BT + j] != 0.f;
            bad |= hv1 && (j + hshift1 >= BTW || j + hshift1 < 0) && wx[int64_t(hg1) * BT + j] != 0.f;
        }
        const int hbase = min(max(hs - xs, 0), G::IW - BTW) & ~1;
        bad |= hv0 && hbase != hs - xs;
    }
    for (int i = tid; i < G::NQ * 4; i += RB_THREADS) {              // V pass: rows 4p .. 4p+3 share one BTV-row window
        const int p = i >> 2, k = i & 3;
        if (4 * p >= th) continue;
        const int r0 = min(oy0 + 4 * p, H - 1), rk = min(oy0 + 4 * p + k, H - 1);
        const int d = loy[rk] - loy[r0];
        for (int t = 0; t < BT; ++t) bad |= (t + d >= BTV || t + d < 0) && wy[int64_t(rk) * BT + t] != 0.f;
        const int y0 = loy[r0] - ys;
        bad |= y0 < 0 || y0 > G::IH - BTV;                           // the quad's window must lie in the staged rows
    }
    if (tid == 0) {
        const int need = max(loy[oy0 + th - 1] + BT, loy[oy0 + ((th - 1) & ~3)] + BT + 2) - ys;
        // rows beyond IH are only harmless when they lie past the image bottom (their weights are zero)
        bad |= need > G::IH && ys + G::IH < H;
    }
    if (bad) atomicExch(flag, 1);
}

template <int BT>
static void rb_prove(const float* tx, const float* ty, int H, int W, int* flag, cudaStream_t st) {
    dim3 grid((W + RB_TW - 1) / RB_TW, (H + RB_TH - 1) / RB_TH);
    rb_prove_kernel<BT><<<grid, RB_THREADS, 0, st>>>(reinterpret_cast<const int*>(tx), tx + W, reinterpret_cast<const int*>(ty), ty + H, H, W, flag);
}

static inline bool rb_ok(int H, int W, int Hm, int Wm, int N) {
    if (H <= 0 || W <= 0 || Hm <= 0 || Wm <= 0 || N <= 0) return false;
    const float rh = (float)Hm / (float)H, rw = (float)Wm / (float)W;
    return rh >= RB_RATIO_MIN && rh <= RB_RATIO_MAX && rw >= RB_RATIO_MIN && rw <= RB_RATIO_MAX &&
           (H + RB_TH - 1) / RB_TH <= 65535 && tmap_encoder() != nullptr;
}

// Window size: the band of A = U D / A^T is at most floor(3 n / nm) + 6 entries; made regular over
// 32 lanes (H pass) or shared by a row pair (V pass) it needs floor(3 n / nm) + 7 (8 when n <= nm)
// (checked over geometries in tests/test_host_cpu.py::test_resize_band_bound).
static inline int rb_band(int H, int W, int Hm, int Wm) {
    const float s = fmaxf((float)H / (float)Hm, (float)W / (float)Wm);
    const int fl = (int)floorf(3.f * s + 1e-3f);
    const int need = fl <= 2 ? 8 : fl + 7;
    return need <= 8 ? 8 : (need <= 10 ? 10 : (need <= 12 ? 12 : 14));
}

static inline RBAxis rb_axis(int n, int nm) { return RBAxis{n, nm, (float)n / (float)nm, (float)nm / (float)n}; }

template <int BT, int DIR, bool RAGGED, int DT = WM_DT_F32>
static int rb_launch(const RBArgs& a, const CUtensorMap& tm, const CUtensorMap& tmm, cudaStream_t st, const char* who) {
    const size_t smem = RBGeom<BT>::smem;
    cudaError_t e = cudaFuncSetAttribute(rb_banded_kernel<BT, DIR, RAGGED, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, who);
    // persistent over planes: ~2 CTAs per SM in total, each tile position walks its share of planes
    const int pos = a.tiles_x * a.tiles_y;
    int gz = (2 * sm_count()) / pos;
    gz = gz < 1 ? 1 : (gz > a.N ? a.N : gz);
    rb_banded_kernel<BT, DIR, RAGGED, DT><<<dim3(a.tiles_x, a.tiles_y, gz), RB_THREADS, smem, st>>>(tm, tmm, a);
    WM_LAUNCH_CHECK(who);
    return WM_OK;
}

}  // namespace wm

using namespace wm;

extern "C" int wm_resize_is_fused(int H, int W, int Hm, int Wm, int N) { return rb_ok(H, W, Hm, Wm, N) ? 1 : 0; }

// floats of table workspace needed by wm_resize_fwd / wm_resize_bwd for this geometry
extern "C" int64_t wm_resize_table_floats(int H, int W, int Hm, int Wm) {
    const int BT = rb_band(H, W, Hm, Wm);
    return 2 * (rb_axis_words(W, BT) + rb_axis_words(H, BT)) + 4;      // + overflow flag
}

extern "C" int wm_resize_tables(float* tables, int H, int W, int Hm, int Wm, int mode, void* stream) {
    WM_REQUIRE(tables, WM_E_NULL, "wm_resize_tables: null workspace");
    WM_REQUIRE(mode == 0 || mode == 1, WM_E_ARG, "wm_resize_tables: mode must be 0 (bilinear) or 1 (bicubic)");
    WM_REQUIRE(H > 0 && W > 0 && Hm > 0 && Wm > 0, WM_E_SHAPE, "wm_resize_tables: bad geometry");
    const int BT = rb_band(H, W, Hm, Wm);
    cudaStream_t st = (cudaStream_t)stream;
    float* fx = tables; float* fy = fx + rb_axis_words(W, BT);
    float* bx = fy + rb_axis_words(H, BT); float* by = bx + rb_axis_words(W, BT);
    const RBAxis ax = rb_axis(W, Wm), ay = rb_axis(H, Hm);
    auto lo = [](float* t) { return reinterpret_cast<int*>(t); };
    if (mode == 0) {
        rb_fwd_tables_kernel<0><<<(W + 127) / 128, 128, 0, st>>>(ax, BT, lo(fx), fx + W);
        rb_fwd_tables_kernel<0><<<(H + 127) / 128, 128, 0, st>>>(ay, BT, lo(fy), fy + H);
    } else {
        rb_fwd_tables_kernel<1><<<(W + 127) / 128, 128, 0, st>>>(ax, BT, lo(fx), fx + W);
        rb_fwd_tables_kernel<1><<<(H + 127) / 128, 128, 0, st>>>(ay, BT, lo(fy), fy + H);
    }
    rb_adj_tables_kernel<<<(W + 127) / 128, 128, 0, st>>>(W, BT, lo(fx), fx + W, lo(bx), bx + W);
    rb_adj_tables_kernel<<<(H + 127) / 128, 128, 0, st>>>(H, BT, lo(fy), fy + H, lo(by), by + H);
    int* flag = reinterpret_cast<int*>(by + rb_axis_words(H, BT));
    cudaMemsetAsync(flag, 0, 4 * sizeof(float), st);
    // prove both directions' tables against the kernel's windows (flag != 0: use wm_interp_fwd twice instead)
    switch (BT) {
        case 8:  rb_prove<8>(fx, fy, H, W, flag, st);  rb_prove<8>(bx, by, H, W, flag, st);  break;
        case 10: rb_prove<10>(fx, fy, H, W, flag, st); rb_prove<10>(bx, by, H, W, flag, st); break;
        case 12: rb_prove<12>(fx, fy, H, W, flag, st); rb_prove<12>(bx, by, H, W, flag, st); break;
        default: rb_prove<14>(fx, fy, H, W, flag, st); rb_prove<14>(bx, by, H, W, flag, st); break;
    }
    WM_LAUNCH_CHECK("wm_resize_tables");
    return WM_OK;
}

// dt: element type of the forward's source planes / of the adjoint's result (WM_DT_F32: the plain kernels)
static int rb_run(int dir, bool ragged, const void* src, int64_t s_sp, int64_t s_sh, void* dst, uint32_t* mask, const float* tables,
                  int N, int H, int W, int Hm, int Wm, const wm_store_epilogue* ep, void* stream, const char* who, int dt = WM_DT_F32) {
    const int BT = rb_band(H, W, Hm, Wm);
    const float* fx = tables; const float* fy = fx + rb_axis_words(W, BT);
    const float* bx = fy + rb_axis_words(H, BT); const float* by = bx + rb_axis_words(W, BT);
    const float* tx = dir == 0 ? fx : bx; const float* ty = dir == 0 ? fy : by;
    RBArgs a{};
    a.src = src; a.s_sp = s_sp; a.s_sh = s_sh; a.dst = dst; a.mask = mask;
    a.lox = reinterpret_cast<const int*>(tx); a.wx = tx + W;
    a.loy = reinterpret_cast<const int*>(ty); a.wy = ty + H;
    a.N = N; a.H = H; a.W = W; a.tiles_x = (W + RB_TW - 1) / RB_TW; a.tiles_y = (H + RB_TH - 1) / RB_TH;
    a.overflow = nullptr;          // proven per geometry at table-build time (rb_prove_kernel), not per launch
    a.ep = dir == 0 ? make_store_ep(ep) : StoreEp{nullptr, 0, 0, 0};
    CUtensorMap tm{}, tmm{};
    int rc = 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool typed_src = dir == 0 && dt != WM_DT_F32;
#define RB_CASE(B)                                                                                                  \
    case B:                                                                                                         \
        if (!ragged)                                                                                                \
            rc = typed_src ? tmap_planes(&tm, dt == WM_DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, \
                                         src, N, H, W, s_sp, s_sh, RBGeom<B>::IW2, RBGeom<B>::IH)                  \
                           : tmap_planes(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, src, N, H, W, s_sp, s_sh, RBGeom<B>::IW, \
                                         RBGeom<B>::IH);                                                            \
        if (!rc && dir == 1 && mask)                                                                                \
            rc = tmap_planes(&tmm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, mask, N, H, 4 * a.tiles_x,                    \
                             int64_t(H) * 4 * a.tiles_x, 4 * a.tiles_x, 12, RBGeom<B>::IH);                         \
        if (rc) {                                                                                                   \
            set_error("%s: cuTensorMapEncodeTiled failed (%d)", who, rc);                                           \
            return WM_E_ARG;                                                                                        \
        }                                                                                                           \
        if (ragged) return dir == 0 ? rb_launch<B, 0, true>(a, tm, tmm, st, who) : rb_launch<B, 1, true>(a, tm, tmm, st, who); \
        if (dt == WM_DT_F16) return dir == 0 ? rb_launch<B, 0, false, WM_DT_F16>(a, tm, tmm, st, who) : rb_launch<B, 1, false, WM_DT_F16>(a, tm, tmm, st, who); \
        if (dt == WM_DT_BF16) return dir == 0 ? rb_launch<B, 0, false, WM_DT_BF16>(a, tm, tmm, st, who) : rb_launch<B, 1, false, WM_DT_BF16>(a, tm, tmm, st, who); \
        return dir == 0 ? rb_launch<B, 0, false>(a, tm, tmm, st, who) : rb_launch<B, 1, false>(a, tm, tmm, st, who);
    switch (BT) {
        RB_CASE(8)
        RB_CASE(10)
        RB_CASE(12)
        RB_CASE(14)
    }
#undef RB_CASE
    set_error("%s: unsupported band %d", who, BT);
    return WM_E_ARG;
}

extern "C" int wm_resize_fwd(const float* x, int64_t x_sp, int64_t x_sh, float* y, int N, int H, int W, int Hm, int Wm,
                             int mode, uint32_t* maskbits, const float* tables, const wm_store_epilogue* ep, void* stream) {
    if (N == 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(x && y && tables, WM_E_NULL, "wm_resize_fwd: null pointer");
    WM_EP_CHECK(ep, "wm_resize_fwd");
    WM_REQUIRE(mode == 0 || mode == 1, WM_E_ARG, "wm_resize_fwd: mode must be 0 (bilinear) or 1 (bicubic)");
    if (N == 0) return WM_OK;
    WM_REQUIRE(rb_ok(H, W, Hm, Wm, N), WM_E_SHAPE,
               "wm_resize_fwd: geometry H=%d W=%d mid=%dx%d N=%d is outside the fused range (ratio %.2f..%.2f); "
               "use wm_interp_fwd twice", H, W, Hm, Wm, N, RB_RATIO_MIN, RB_RATIO_MAX);
    WM_REQUIRE(!maskbits || aligned(maskbits, 16), WM_E_ALIGN, "wm_resize_fwd: maskbits must be 16-byte aligned");
    // rows on 16-byte boundaries: TMA; otherwise (W % 4 != 0, odd strides) the cp.async-fed instantiation
    const bool ragged = !(W % 4 == 0 && tmap_ok(x, x_sp, x_sh, 4) && aligned(y, 16));
    if (ragged) WM_EP_REJECT(ep, "wm_resize_fwd (rows not 16-byte aligned)");
    return rb_run(0, ragged, x, x_sp, x_sh, y, maskbits, tables, N, H, W, Hm, Wm, ep, stream, "wm_resize_fwd");
}

extern "C" int wm_resize_bwd(const float* gy, const uint32_t* maskbits, float* gx, int N, int H, int W, int Hm, int Wm,
                             int mode, const float* tables, void* stream) {
    if (N == 0) return WM_OK;      // empty work: nothing to validate or launch
    WM_REQUIRE(gy && gx && tables, WM_E_NULL, "wm_resize_bwd: null pointer");
    WM_REQUIRE(mode == 0 || mode == 1, WM_E_ARG, "wm_resize_bwd: mode must be 0 (bilinear) or 1 (bicubic)");
    if (N == 0) return WM_OK;
    WM_REQUIRE(rb_ok(H, W, Hm, Wm, N), WM_E_SHAPE,
               "wm_resize_bwd: geometry H=%d W=%d mid=%dx%d N=%d is outside the fused range; use wm_interp_bwd twice",
               H, W, Hm, Wm, N);
    WM_REQUIRE(!maskbits || aligned(maskbits, 16), WM_E_ALIGN, "wm_resize_bwd: maskbits must be 16-byte aligned");
    const bool ragged = !(W % 4 == 0 && aligned(gy, 16) && aligned(gx, 16));
    return rb_run(1, ragged, gy, int64_t(H) * W, W, gx, const_cast<uint32_t*>(maskbits), tables, N, H, W, Hm, Wm, nullptr, stream, "wm_resize_bwd");
}

// Typed planes (include/wm_attack.h): float16 / bfloat16 source planes staged as they are (forward), or the adjoint's
// result stored in that type.  Rows on 16-byte boundaries only; no store epilogue.
extern "C" int wm_resize_fwd_typed(const void* x, int x_dtype, int64_t x_sp, int64_t x_sh, float* y, int N, int H, int W, int Hm, int Wm,
                                   int mode, uint32_t* maskbits, const float* tables, void* stream) {
    if (N == 0) return WM_OK;
    if (x_dtype == WM_DT_F32)
        return wm_resize_fwd(reinterpret_cast<const float*>(x), x_sp, x_sh, y, N, H, W, Hm, Wm, mode, maskbits, tables, nullptr, stream);
    WM_REQUIRE(x && y && tables, WM_E_NULL, "wm_resize_fwd_typed: null pointer");
    WM_REQUIRE(x_dtype == WM_DT_F16 || x_dtype == WM_DT_BF16, WM_E_ARG, "wm_resize_fwd_typed: unknown element type %d", x_dtype);
    WM_REQUIRE(mode == 0 || mode == 1, WM_E_ARG, "wm_resize_fwd_typed: mode must be 0 (bilinear) or 1 (bicubic)");
    WM_REQUIRE(rb_ok(H, W, Hm, Wm, N), WM_E_SHAPE, "wm_resize_fwd_typed: geometry H=%d W=%d mid=%dx%d N=%d is outside the fused range", H, W, Hm, Wm, N);
    WM_REQUIRE(!maskbits || aligned(maskbits, 16), WM_E_ALIGN, "wm_resize_fwd_typed: maskbits must be 16-byte aligned");
    WM_REQUIRE(W % 4 == 0 && tmap_ok(x, x_sp, x_sh, 2) && aligned(y, 16), WM_E_ALIGN,
               "wm_resize_fwd_typed: 2-byte planes need rows on 16-byte boundaries (W %% 8 == 0, aligned strides); convert to float32 otherwise");
    return rb_run(0, false, x, x_sp, x_sh, y, maskbits, tables, N, H, W, Hm, Wm, nullptr, stream, "wm_resize_fwd_typed", x_dtype);
}

extern "C" int wm_resize_bwd_typed(const float* gy, const uint32_t* maskbits, void* gx, int gx_dtype, int N, int H, int W, int Hm, int Wm,
                                   int mode, const float* tables, void* stream) {
    if (N == 0) return WM_OK;
    if (gx_dtype == WM_DT_F32) return wm_resize_bwd(gy, maskbits, reinterpret_cast<float*>(gx), N, H, W, Hm, Wm, mode, tables, stream);
    WM_REQUIRE(gy && gx && tables, WM_E_NULL, "wm_resize_bwd_typed: null pointer");
    WM_REQUIRE(gx_dtype == WM_DT_F16 || gx_dtype == WM_DT_BF16, WM_E_ARG, "wm_resize_bwd_typed: unknown element type %d", gx_dtype);
    WM_REQUIRE(mode == 0 || mode == 1, WM_E_ARG, "wm_resize_bwd_typed: mode must be 0 (bilinear) or 1 (bicubic)");
    WM_REQUIRE(rb_ok(H, W, Hm, Wm, N), WM_E_SHAPE, "wm_resize_bwd_typed: geometry H=%d W=%d mid=%dx%d N=%d is outside the fused range", H, W, Hm, Wm, N);
    WM_REQUIRE(!maskbits || aligned(maskbits, 16), WM_E_ALIGN, "wm_resize_bwd_typed: maskbits must be 16-byte aligned");
    WM_REQUIRE(W % 4 == 0 && aligned(gy, 16) && aligned(gx, 8), WM_E_ALIGN, "wm_resize_bwd_typed: needs W %% 4 == 0 and aligned planes");
    return rb_run(1, false, gy, int64_t(H) * W, W, gx, const_cast<uint32_t*>(maskbits), tables, N, H, W, Hm, Wm, nullptr, stream,
                  "wm_resize_bwd_typed", gx_dtype);
}
