// Fused Resize round trip:  y = clamp( up( down(x) ), 0, 1 )   and its exact adjoint.
//
// Replaces Resize.forward (noise_layers/resize.py:38-53): F.interpolate to int(r*H) x int(r*W),
// F.interpolate back to H x W, clamp — three full-tensor passes plus the materialised mid image
// (24 + 24 r^2 B/px) — by ONE kernel per direction that reads x once and writes y once (24 B/px;
// backward: gy + 1 bit/value mask in, gx out).
//
// Both interpolations are separable 4-tap gathers along each axis with F.interpolate's
// align_corners=False coordinates (ATen upsample_bicubic2d A=-0.75 / upsample_bilinear2d), taps
// index-clamped at the borders == 4 CONSECUTIVE positions of the edge-replicated signal.  Along
// one axis the round trip is therefore a STREAM: walking the input positions p in order, a mid
// sample q is complete when its last tap (position end_p[q]) has arrived, an output o when its
// last mid tap (end_q[o]) has; both windows are the last 4 values seen, held in registers.
// All lanes of a warp walk the same positions (the schedule depends only on the tile), so the
// control flow is uniform:
//   H pass: warp = 16-output column segment, lane = image row (3 rows per lane), in -> tmp
//   V pass: warp = 8-output row segment,    lane = image column (4 per lane),   tmp -> y (+mask)
// Every shared-memory value is read once per pass (a tap-table gather would read each 4 times and
// saturate the LDS pipe before HBM).  The adjoint runs the same streams backwards as scatters
// into 4-wide accumulator windows: a deterministic, atomic-free transpose.
#include "wm_common.cuh"

namespace wm {

constexpr int RF_TW = 128, RF_TH = 64, RF_THREADS = 256;
constexpr int RF_IWMAX = RF_TW + 18, RF_IHMAX = RF_TH + 18;          // staged region bounds
constexpr int RF_IP = RF_IWMAX | 1, RF_TP = RF_TW + 1;               // odd pitches: lanes = rows is conflict-free
constexpr int RF_NQH = 304, RF_NQV = 164;                            // mid positions per tile axis (ratio <= 2.2)
constexpr float RF_RATIO_MIN = 0.45f, RF_RATIO_MAX = 2.2f;
constexpr int RF_HSEG = 16, RF_HL = 3, RF_VSEG = 8, RF_VL = 4, RF_LROWS = 32 * RF_HL;

__device__ __forceinline__ float rf_cubic1(float x) { const float A = -0.75f; return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float rf_cubic2(float x) { const float A = -0.75f; return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

// Taps of output index o: positions i0-1 .. i0+2 of the edge-replicated source, weights w.
// Same fp32 coordinate arithmetic as ATen (area_pixel_compute_source_index, align_corners=false).
template <int MODE>
__device__ __forceinline__ void rf_tap(float scale, int o, int n_in, int& i0, float4& w) {
    float rho = scale * (o + 0.5f) - 0.5f;
    if (MODE == 0) {
        rho = fmaxf(rho, 0.f);
        i0 = min(int(rho), n_in - 1);
        const float l1 = fminf(fmaxf(rho - i0, 0.f), 1.f);
        w = make_float4(0.f, 1.f - l1, l1, 0.f);
    } else {
        const float fl = floorf(rho);
        i0 = int(fl);
        const float t = rho - fl;
        w = make_float4(rf_cubic2(t + 1.f), rf_cubic1(t), rf_cubic1(1.f - t), rf_cubic2(2.f - t));
    }
}

struct RFAxis { int n, nm; float sd, su; };       // size, mid size, down scale n/nm, up scale nm/n

struct RFArgs {
    const float* x; int64_t x_sp, x_sh;           // forward input / backward cotangent (dense)
    float* y;                                     // forward output / backward gx
    uint32_t* mask; int mask_wpr;                 // clamp pass-through bits, words per row
    int N; RFAxis ax, ay;                         // ax: columns (W), ay: rows (H)
    int tiles_x, tiles_y;
};

// Per-axis stream tables of one tile (shared memory)
template <int NOUT, int NQMAX>
struct RFTables {
    float4 wu[NOUT];     // up-taps of the tile's outputs
    int endq[NOUT];      // local mid position of the last up-tap
    float4 wd[NQMAX];    // down-taps of the mid positions
    int endp[NQMAX];     // local input position of the last down-tap
    int nq, np;          // mid positions, input positions of the tile
    int p_a;             // absolute (unclamped) input position of local position 0
    int pc_a, pc_n;      // first loaded (clamped) input index, number of loaded indices
};

// Builds the tables for outputs [o0, o0 + nout) of an axis.  Called by all threads of the CTA.
template <int MODE, int NOUT, int NQMAX>
__device__ __forceinline__ void rf_build_tables(RFTables<NOUT, NQMAX>& T, const RFAxis ax, int o0, int nout, int* iu0_s, int* id0_s) {
    const int t = threadIdx.x;
    if (t < NOUT) {
        int i0; float4 w;
        rf_tap<MODE>(ax.su, min(o0 + min(t, nout - 1), ax.n - 1), ax.nm, i0, w);
        T.wu[t] = w; iu0_s[t] = i0;
    }
    __syncthreads();
    const int q_a = iu0_s[0] - 1, q_b = iu0_s[nout - 1] + 2;
    const int nq = min(q_b - q_a + 1, NQMAX);
    for (int i = t; i < nq; i += RF_THREADS) {
        int i0; float4 w;
        rf_tap<MODE>(ax.sd, min(max(q_a + i, 0), ax.nm - 1), ax.n, i0, w);
        T.wd[i] = w; id0_s[i] = i0;
    }
    __syncthreads();
    const int p_a = id0_s[0] - 1, p_b = id0_s[nq - 1] + 2;
    if (t < NOUT) T.endq[t] = iu0_s[t] + 2 - q_a;
    for (int i = t; i < nq; i += RF_THREADS) T.endp[i] = id0_s[i] + 2 - p_a;
    if (t == 0) {
        T.nq = nq; T.np = p_b - p_a + 1; T.p_a = p_a;
        const int ca = min(max(p_a, 0), ax.n - 1), cb = min(max(p_b, 0), ax.n - 1);
        T.pc_a = ca; T.pc_n = cb - ca + 1;
    }
    __syncthreads();
}

// One stream segment: outputs [o_s, o_e) of the tile axis for the L lines held by this thread.
//   value of line j at loaded index pi :  lbase[pi * LPOS + j * LLINE]   (immediate line offsets)
//   emit(o, v[L])                      <- output o (tile-local) of every line
// The input window is rotated statically (the position loop is unrolled by 4, so slot = p & 3 is a
// compile-time index); the mid window shifts.
template <int L, int LLINE, int LPOS, int NOUT, int NQMAX, typename Emit>
__device__ __forceinline__ void rf_stream(const RFTables<NOUT, NQMAX>& T, int n_axis, int o_s, int o_e,
                                          const float* __restrict__ lbase, Emit emit) {
    float a[L][4], m[L][4];
#pragma unroll
    for (int j = 0; j < L; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) { a[j][k] = 0.f; m[j][k] = 0.f; }
    int q = T.endq[o_s] - 3;                 // first mid position of the segment's first window
    int p = T.endp[q] - 3;                   // first input position of that mid's window
    int o = o_s;
    int eq = T.endp[q], eo = T.endq[o];
    const int nq = T.nq, np = T.np, pc_a = T.pc_a;
    int pabs = T.p_a + p;                    // absolute (unclamped) input position
    bool done = false;
    while (!done) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            if (!done) {
                const float* lp = lbase + (min(max(pabs, 0), n_axis - 1) - pc_a) * LPOS;
#pragma unroll
                for (int j = 0; j < L; ++j) a[j][s] = lp[j * LLINE];
                while (eq == p) {
                    const float4 w = T.wd[q];
#pragma unroll
                    for (int j = 0; j < L; ++j) {
                        const float v = fmaf(w.w, a[j][s], fmaf(w.z, a[j][(s + 3) & 3], fmaf(w.y, a[j][(s + 2) & 3], w.x * a[j][(s + 1) & 3])));
                        m[j][0] = m[j][1]; m[j][1] = m[j][2]; m[j][2] = m[j][3]; m[j][3] = v;
                    }
                    while (eo == q) {
                        const float4 u = T.wu[o];
                        float v[L];
#pragma unroll
                        for (int j = 0; j < L; ++j)
                            v[j] = fmaf(u.w, m[j][3], fmaf(u.z, m[j][2], fmaf(u.y, m[j][1], u.x * m[j][0])));
                        emit(o, v);
                        ++o;
                        if (o < o_e) eo = T.endq[o]; else { eo = -2; done = true; }
                    }
                    ++q;
                    eq = (q < nq && !done) ? T.endp[q] : -2;
                }
                ++p; ++pabs;
                if (p >= np) done = true;    // cannot spin if a table was truncated
            }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(RF_THREADS, 2) resize_fused_fwd_kernel(const RFArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using TH_t = RFTables<RF_TW, RF_NQH>;
    using TV_t = RFTables<RF_TH, RF_NQV>;
    TH_t& Tx = *reinterpret_cast<TH_t*>(smem_raw);
    TV_t& Ty = *reinterpret_cast<TV_t*>(smem_raw + sizeof(TH_t));
    float* in = reinterpret_cast<float*>(smem_raw + sizeof(TH_t) + sizeof(TV_t));   // [RF_IHMAX][RF_IP]
    float* tmp = in + RF_IHMAX * RF_IP;                                             // [RF_LROWS][RF_TP]
    int* scratch = reinterpret_cast<int*>(tmp);   // table construction scratch (tmp is not live yet)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ox0 = blockIdx.x * RF_TW, oy0 = blockIdx.y * RF_TH;
    const int tw = min(RF_TW, a.ax.n - ox0), th = min(RF_TH, a.ay.n - oy0);

    // the stream tables depend on the tile position only: built once, reused for every plane
    rf_build_tables<MODE>(Tx, a.ax, ox0, tw, scratch, scratch + RF_TW);
    rf_build_tables<MODE>(Ty, a.ay, oy0, th, scratch, scratch + RF_TW);
    const int IW = Tx.pc_n, IH = Ty.pc_n;
    const uint32_t in_s = (uint32_t)__cvta_generic_to_shared(in);

    for (int n = blockIdx.z; n < a.N; n += gridDim.z) {
        // stage the source region: 4-byte cp.async (LDGSTS), no register staging, all copies in flight
        {
            const float* row = a.x + int64_t(n) * a.x_sp + int64_t(Ty.pc_a + warp) * a.x_sh + Tx.pc_a + lane;
            uint32_t drow = in_s + 4u * (warp * RF_IP + lane);
            for (int r = warp; r < IH; r += RF_THREADS / 32) {
#pragma unroll
                for (int i = 0; i < (RF_IWMAX + 31) / 32; ++i)
                    if (lane + 32 * i < IW)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(drow + 128u * i), "l"(row + 32 * i) : "memory");
                row += int64_t(RF_THREADS / 32) * a.x_sh;
                drow += 4u * (RF_THREADS / 32) * RF_IP;
            }
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();

        // H pass: in[IH][IW] -> tmp[IH][tw]; lane = row (+32, +64), warp = 16-column segment.
        // Lines beyond IH run on whatever shared memory holds there and land in spare tmp rows.
        {
            const int o_s = warp * RF_HSEG, o_e = min(o_s + RF_HSEG, tw);
            if (o_s < o_e) {
                float* trow = tmp + lane * RF_TP;
                rf_stream<RF_HL, 32 * RF_IP, 1>(Tx, a.ax.n, o_s, o_e, in + lane * RF_IP, [&](int o, const float (&v)[RF_HL]) {
#pragma unroll
                    for (int j = 0; j < RF_HL; ++j) trow[j * 32 * RF_TP + o] = v[j];
                });
            }
        }
        __syncthreads();

        // V pass: tmp[IH][tw] -> y (clamp, mask); lane = column (+32, +64, +96), warp = 8-row segment
        {
            const int o_s = warp * RF_VSEG, o_e = min(o_s + RF_VSEG, th);
            if (o_s < o_e) {
                float* dst = a.y + (int64_t(n) * a.ay.n + oy0) * a.ax.n + ox0 + lane;
                uint32_t* mdst = a.mask ? a.mask + (int64_t(n) * a.ay.n + oy0) * a.mask_wpr + (ox0 >> 5) + lane : nullptr;
                const int nwords = (tw + 31) >> 5;
                bool okc[RF_VL];
#pragma unroll
                for (int j = 0; j < RF_VL; ++j) okc[j] = lane + 32 * j < tw;
                rf_stream<RF_VL, 32, RF_TP>(Ty, a.ay.n, o_s, o_e, tmp + lane, [&](int o, const float (&v)[RF_VL]) {
                    float* drow = dst + int64_t(o) * a.ax.n;
                    unsigned word = 0;
#pragma unroll
                    for (int j = 0; j < RF_VL; ++j) {
                        const float c = __saturatef(v[j]);
                        if (okc[j]) drow[32 * j] = c;
                        if (mdst) {     // 0 <= v <= 1  <=>  saturate(v) == v  (false for NaN)
                            const unsigned bits = __ballot_sync(0xffffffffu, okc[j] && c == v[j]);
                            word = lane == j ? bits : word;
                        }
                    }
                    if (mdst && lane < nwords) mdst[int64_t(o) * a.mask_wpr] = word;
                });
            }
        }
        // the next plane's staging writes `in` (dead since the barrier above) and its barrier orders
        // the next H pass (which writes tmp) after every warp's V pass
    }
}

// =============================================================================================
// adjoint:  gx = D_v^T U_v^T (gy .* mask) U_h D_h      (same streams, run as scatters)
// =============================================================================================
// Along one axis, for the gx indices [i_a, i_b] of a tile (or of a warp's segment):
//   positions p with clamp(p) in [i_a, i_b]  <- mid positions q whose down-window touches them
//   <- outputs o whose up-window touches those q.  A line walks q upwards; outputs whose window
// STARTS at q scatter u[o][k] * g[o] into a 4-wide mid accumulator window, the finished front
// element gm[q] scatters d[q][k] * gm into a 4-wide input accumulator window, whose front
// elements leave as gx (positions folded onto the clamped border indices are summed first).
struct RFAdjRange {
    int o_lo, o_hi;          // outputs that contribute
    int q_first, q_last;     // mid positions walked (the window starts of o_lo .. window end of o_hi)
    int q_lo, q_hi;          // mid positions that scatter into [i_a, i_b]
    int p_first, p_last;     // input positions emitted
};

template <int MODE>
__device__ __forceinline__ int rf_id0(const RFAxis ax, int q) {       // first tap position - of mid position q
    int i0; float4 w;
    rf_tap<MODE>(ax.sd, min(max(q, 0), ax.nm - 1), ax.n, i0, w);
    return i0;
}
template <int MODE>
__device__ __forceinline__ int rf_iu0(const RFAxis ax, int o) {
    int i0; float4 w;
    rf_tap<MODE>(ax.su, o, ax.nm, i0, w);
    return i0;
}

// Executed by one thread.  All searches start from the analytic inverse and move a few steps.
template <int MODE>
__device__ void rf_adj_range(const RFAxis ax, int i_a, int i_b, RFAdjRange& R) {
    const int BIG = 1 << 28;
    const int P_a = i_a == 0 ? -BIG : i_a, P_b = i_b == ax.n - 1 ? BIG : i_b;
    const int Qmin = rf_iu0<MODE>(ax, 0) - 1, Qmax = rf_iu0<MODE>(ax, ax.n - 1) + 2;
    // q_lo: smallest q with id0(q) + 2 >= P_a
    int q = i_a == 0 ? Qmin : min(max(int(floorf((i_a - 1.5f) * ax.su - 0.5f)) - 2, Qmin), Qmax);
    for (int it = 0; it < 64 && q > Qmin && rf_id0<MODE>(ax, q - 1) + 2 >= P_a; ++it) --q;
    for (int it = 0; it < 64 && q < Qmax && rf_id0<MODE>(ax, q) + 2 < P_a; ++it) ++q;
    R.q_lo = q;
    // q_hi: largest q with id0(q) - 1 <= P_b
    q = i_b == ax.n - 1 ? Qmax : min(max(int(floorf((i_b + 1.5f) * ax.su - 0.5f)) + 2, Qmin), Qmax);
    for (int it = 0; it < 64 && q < Qmax && rf_id0<MODE>(ax, q + 1) - 1 <= P_b; ++it) ++q;
    for (int it = 0; it < 64 && q > Qmin && rf_id0<MODE>(ax, q) - 1 > P_b; ++it) --q;
    R.q_hi = q;
    // o_lo: smallest o with iu0(o) + 2 >= q_lo ;  o_hi: largest o with iu0(o) - 1 <= q_hi
    int o = min(max(int(floorf((R.q_lo - 1.5f) * ax.sd - 0.5f)) - 2, 0), ax.n - 1);
    for (int it = 0; it < 64 && o > 0 && rf_iu0<MODE>(ax, o - 1) + 2 >= R.q_lo; ++it) --o;
    for (int it = 0; it < 64 && o < ax.n - 1 && rf_iu0<MODE>(ax, o) + 2 < R.q_lo; ++it) ++o;
    R.o_lo = o;
    o = min(max(int(floorf((R.q_hi + 1.5f) * ax.sd - 0.5f)) + 2, 0), ax.n - 1);
    for (int it = 0; it < 64 && o < ax.n - 1 && rf_iu0<MODE>(ax, o + 1) - 1 <= R.q_hi; ++it) ++o;
    for (int it = 0; it < 64 && o > 0 && rf_iu0<MODE>(ax, o) - 1 > R.q_hi; ++it) --o;
    R.o_hi = o;
    R.q_first = min(R.q_lo, rf_iu0<MODE>(ax, R.o_lo) - 1);
    R.q_last = max(R.q_hi, rf_iu0<MODE>(ax, R.o_hi) + 2);
    R.p_first = rf_id0<MODE>(ax, R.q_lo) - 1;
    R.p_last = rf_id0<MODE>(ax, R.q_hi) + 2;
}

template <int NOMAX, int NQMAX>
struct RFAdjTables {
    float4 wu[NOMAX]; int startq[NOMAX];      // per output o - o_lo: up taps, absolute first mid position
    float4 wd[NQMAX]; int startp[NQMAX];      // per mid position q - q_first: down taps, absolute first input position
    RFAdjRange tile;
    RFAdjRange seg[RF_THREADS / 32];
};

// Builds the tile tables and the per-warp segment ranges.  Called by all threads.
template <int MODE, int NOMAX, int NQMAX>
__device__ __forceinline__ void rf_build_adj(RFAdjTables<NOMAX, NQMAX>& T, const RFAxis ax, int i0, int ni, int seg_len) {
    const int t = threadIdx.x;
    if (t == 0) rf_adj_range<MODE>(ax, i0, i0 + ni - 1, T.tile);
    if ((t & 31) == 1) {                  // one lane per warp: the warp's own segment
        const int w = t >> 5, s_a = i0 + w * seg_len, s_b = min(s_a + seg_len, i0 + ni) - 1;
        if (s_a <= s_b) rf_adj_range<MODE>(ax, s_a, s_b, T.seg[w]);
        else T.seg[w].o_lo = 1, T.seg[w].o_hi = 0, T.seg[w].q_first = 1, T.seg[w].q_last = 0;
    }
    __syncthreads();
    const RFAdjRange R = T.tile;
    const int no = min(R.o_hi - R.o_lo + 1, NOMAX), nq = min(R.q_last - R.q_first + 1, NQMAX);
    for (int i = t; i < no; i += RF_THREADS) {
        int b; float4 w;
        rf_tap<MODE>(ax.su, R.o_lo + i, ax.nm, b, w);
        T.wu[i] = w; T.startq[i] = b - 1;
    }
    for (int i = t; i < nq; i += RF_THREADS) {
        int b; float4 w;
        rf_tap<MODE>(ax.sd, min(max(R.q_first + i, 0), ax.nm - 1), ax.n, b, w);
        T.wd[i] = w; T.startp[i] = b - 1;
    }
    __syncthreads();
}

// One adjoint stream segment for L lines.
//   cotangent of line j at output o (absolute): gbase[(o - o_base) * GPOS + j * GLINE]
//   emit(c, v[L]) <- gradient of input index c (absolute, inside [i_a, i_b])
template <int L, int GLINE, int GPOS, int NOMAX, int NQMAX, typename Emit>
__device__ __forceinline__ void rf_adj_stream(const RFAdjTables<NOMAX, NQMAX>& T, const RFAdjRange R, int n_axis, int i_a, int i_b,
                                              const float* __restrict__ gbase, int o_base, Emit emit) {
    float M[L][4], A[L][4], carry[L];
#pragma unroll
    for (int j = 0; j < L; ++j) {
        carry[j] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) { M[j][k] = 0.f; A[j][k] = 0.f; }
    }
    const int o_t = T.tile.o_lo, q_t = T.tile.q_first;
    int o = R.o_lo, pb = R.p_first;
    int so = o <= R.o_hi ? T.startq[o - o_t] : (1 << 29);
    auto pop = [&]() {                        // the front input accumulator is final: fold + emit
        const int c = min(max(pb, 0), n_axis - 1), cn = min(max(pb + 1, 0), n_axis - 1);
#pragma unroll
        for (int j = 0; j < L; ++j) { carry[j] += A[j][0]; A[j][0] = A[j][1]; A[j][1] = A[j][2]; A[j][2] = A[j][3]; A[j][3] = 0.f; }
        if (cn != c || pb == R.p_last) {
            if (c >= i_a && c <= i_b) emit(c, carry);
#pragma unroll
            for (int j = 0; j < L; ++j) carry[j] = 0.f;
        }
        ++pb;
    };
    for (int q = R.q_first; q <= R.q_last; ++q) {
        while (so == q) {
            const float4 u = T.wu[o - o_t];
            const float* gp = gbase + (o - o_base) * GPOS;
#pragma unroll
            for (int j = 0; j < L; ++j) {
                const float g = gp[j * GLINE];
                M[j][0] = fmaf(u.x, g, M[j][0]); M[j][1] = fmaf(u.y, g, M[j][1]);
                M[j][2] = fmaf(u.z, g, M[j][2]); M[j][3] = fmaf(u.w, g, M[j][3]);
            }
            ++o;
            so = o <= R.o_hi ? T.startq[o - o_t] : (1 << 29);
        }
        if (q >= R.q_lo && q <= R.q_hi) {
            const int sp = T.startp[q - q_t];
            while (pb < sp && pb <= R.p_last) pop();   // second test: cannot spin on a corrupt table
            const float4 d = T.wd[q - q_t];
#pragma unroll
            for (int j = 0; j < L; ++j) {
                const float gm = M[j][0];
                A[j][0] = fmaf(d.x, gm, A[j][0]); A[j][1] = fmaf(d.y, gm, A[j][1]);
                A[j][2] = fmaf(d.z, gm, A[j][2]); A[j][3] = fmaf(d.w, gm, A[j][3]);
            }
        }
#pragma unroll
        for (int j = 0; j < L; ++j) { M[j][0] = M[j][1]; M[j][1] = M[j][2]; M[j][2] = M[j][3]; M[j][3] = 0.f; }
    }
    while (pb <= R.p_last) pop();
}

constexpr int RF_GWMAX = RF_IWMAX, RF_GHMAX = RF_IHMAX;

template <int MODE>
__global__ void __launch_bounds__(RF_THREADS, 2) resize_fused_bwd_kernel(const RFArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using TH_t = RFAdjTables<RF_GWMAX, RF_NQH>;
    using TV_t = RFAdjTables<RF_GHMAX, RF_NQV>;
    TH_t& Tx = *reinterpret_cast<TH_t*>(smem_raw);
    TV_t& Ty = *reinterpret_cast<TV_t*>(smem_raw + sizeof(TH_t));
    float* G = reinterpret_cast<float*>(smem_raw + sizeof(TH_t) + sizeof(TV_t));    // [RF_GHMAX][RF_IP] masked cotangent
    float* tmp = G + RF_GHMAX * RF_IP;                                              // [RF_LROWS][RF_TP]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ix0 = blockIdx.x * RF_TW, iy0 = blockIdx.y * RF_TH;
    const int tw = min(RF_TW, a.ax.n - ix0), th = min(RF_TH, a.ay.n - iy0);

    rf_build_adj<MODE>(Tx, a.ax, ix0, tw, RF_HSEG);
    rf_build_adj<MODE>(Ty, a.ay, iy0, th, RF_VSEG);
    const int ox_lo = Tx.tile.o_lo, GW = min(Tx.tile.o_hi - ox_lo + 1, RF_GWMAX);
    const int oy_lo = Ty.tile.o_lo, GH = min(Ty.tile.o_hi - oy_lo + 1, RF_GHMAX);

    for (int n = blockIdx.z; n < a.N; n += gridDim.z) {
        // stage gy .* mask for the output region the tile depends on
        {
            const float* gsrc = a.x + (int64_t(n) * a.ay.n + oy_lo) * a.ax.n + ox_lo;
            const uint32_t* msrc = a.mask ? a.mask + (int64_t(n) * a.ay.n + oy_lo) * a.mask_wpr : nullptr;
            for (int r = warp; r < GH; r += RF_THREADS / 32) {
                const float* row = gsrc + int64_t(r) * a.ax.n;
                float v[(RF_GWMAX + 31) / 32];
#pragma unroll
                for (int i = 0; i < (RF_GWMAX + 31) / 32; ++i) {
                    const int c = lane + 32 * i;
                    v[i] = c < GW ? __ldg(row + c) : 0.f;
                }
                if (msrc) {
#pragma unroll
                    for (int i = 0; i < (RF_GWMAX + 31) / 32; ++i) {
                        const int ox = ox_lo + lane + 32 * i;
                        if (lane + 32 * i < GW) {
                            const uint32_t w = __ldg(msrc + int64_t(r) * a.mask_wpr + (ox >> 5));
                            v[i] = (w >> (ox & 31)) & 1u ? v[i] : 0.f;
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < (RF_GWMAX + 31) / 32; ++i)
                    if (lane + 32 * i < GW) G[r * RF_IP + lane + 32 * i] = v[i];
            }
        }
        __syncthreads();

        // H adjoint: G[GH][GW] -> tmp[GH][tw]; lane = output row, warp = 16 gx columns
        {
            const RFAdjRange R = Tx.seg[warp];
            const int s_a = ix0 + warp * RF_HSEG, s_b = min(s_a + RF_HSEG, ix0 + tw) - 1;
            if (s_a <= s_b) {
                float* trow = tmp + lane * RF_TP - ix0;
                rf_adj_stream<RF_HL, 32 * RF_IP, 1>(Tx, R, a.ax.n, s_a, s_b, G + lane * RF_IP, ox_lo, [&](int c, const float (&v)[RF_HL]) {
#pragma unroll
                    for (int j = 0; j < RF_HL; ++j) trow[j * 32 * RF_TP + c] = v[j];
                });
            }
        }
        __syncthreads();

        // V adjoint: tmp[GH][tw] -> gx; lane = column, warp = 8 gx rows
        {
            const RFAdjRange R = Ty.seg[warp];
            const int s_a = iy0 + warp * RF_VSEG, s_b = min(s_a + RF_VSEG, iy0 + th) - 1;
            if (s_a <= s_b) {
                float* dst = a.y + int64_t(n) * a.ay.n * a.ax.n + ix0 + lane;
                bool okc[RF_VL];
#pragma unroll
                for (int j = 0; j < RF_VL; ++j) okc[j] = lane + 32 * j < tw;
                rf_adj_stream<RF_VL, 32, RF_TP>(Ty, R, a.ay.n, s_a, s_b, tmp + lane, oy_lo, [&](int c, const float (&v)[RF_VL]) {
                    float* drow = dst + int64_t(c) * a.ax.n;
#pragma unroll
                    for (int j = 0; j < RF_VL; ++j)
                        if (okc[j]) drow[32 * j] = v[j];
                });
            }
        }
    }
}

static inline bool rf_ok(int H, int W, int Hm, int Wm, int N) {
    if (H <= 0 || W <= 0 || Hm <= 0 || Wm <= 0 || N <= 0) return false;
    const float rh = (float)Hm / (float)H, rw = (float)Wm / (float)W;
    return rh >= RF_RATIO_MIN && rh <= RF_RATIO_MAX && rw >= RF_RATIO_MIN && rw <= RF_RATIO_MAX &&
           (H + RF_TH - 1) / RF_TH <= 65535;
}

static inline size_t rf_smem() {
    return sizeof(RFTables<RF_TW, RF_NQH>) + sizeof(RFTables<RF_TH, RF_NQV>) +
           sizeof(float) * (size_t(RF_IHMAX) * RF_IP + size_t(RF_LROWS) * RF_TP);
}

static inline size_t rf_smem_bwd() {
    return sizeof(RFAdjTables<RF_GWMAX, RF_NQH>) + sizeof(RFAdjTables<RF_GHMAX, RF_NQV>) +
           sizeof(float) * (size_t(RF_GHMAX) * RF_IP + size_t(RF_LROWS) * RF_TP);
}

}  // namespace wm

using namespace wm;

extern "C" int wm_resize_bwd(const float* gy, const uint32_t* maskbits, float* gx, int N, int H, int W, int Hm, int Wm,
                             int mode, void* stream) {
    WM_REQUIRE(gy && gx, WM_E_NULL, "wm_resize_bwd: null pointer");
    WM_REQUIRE(mode == 0 || mode == 1, WM_E_ARG, "wm_resize_bwd: mode must be 0 (bilinear) or 1 (bicubic)");
    if (N == 0) return WM_OK;
    WM_REQUIRE(rf_ok(H, W, Hm, Wm, N), WM_E_SHAPE,
               "wm_resize_bwd: geometry H=%d W=%d mid=%dx%d N=%d is outside the fused range; use wm_interp_bwd twice",
               H, W, Hm, Wm, N);
    RFArgs a{};
    a.x = gy; a.x_sp = int64_t(H) * W; a.x_sh = W; a.y = gx; a.mask = const_cast<uint32_t*>(maskbits);
    a.mask_wpr = (W + 31) / 32; a.N = N;
    a.ax = RFAxis{W, Wm, (float)W / (float)Wm, (float)Wm / (float)W};
    a.ay = RFAxis{H, Hm, (float)H / (float)Hm, (float)Hm / (float)H};
    a.tiles_x = (W + RF_TW - 1) / RF_TW; a.tiles_y = (H + RF_TH - 1) / RF_TH;
    const size_t smem = rf_smem_bwd();
    auto kern = mode == 0 ? resize_fused_bwd_kernel<0> : resize_fused_bwd_kernel<1>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "wm_resize_bwd");
    const int pos = a.tiles_x * a.tiles_y;
    int gz = (2 * sm_count()) / pos;
    gz = gz < 1 ? 1 : (gz > N ? N : gz);
    kern<<<dim3(a.tiles_x, a.tiles_y, gz), RF_THREADS, smem, (cudaStream_t)stream>>>(a);
    WM_LAUNCH_CHECK("wm_resize_bwd");
    return WM_OK;
}

extern "C" int wm_resize_is_fused(int H, int W, int Hm, int Wm, int N) { return rf_ok(H, W, Hm, Wm, N) ? 1 : 0; }

extern "C" int wm_resize_fwd(const float* x, int64_t x_sp, int64_t x_sh, float* y, int N, int H, int W, int Hm, int Wm,
                             int mode, uint32_t* maskbits, void* stream) {
    WM_REQUIRE(x && y, WM_E_NULL, "wm_resize_fwd: null pointer");
    WM_REQUIRE(mode == 0 || mode == 1, WM_E_ARG, "wm_resize_fwd: mode must be 0 (bilinear) or 1 (bicubic)");
    if (N == 0) return WM_OK;
    WM_REQUIRE(rf_ok(H, W, Hm, Wm, N), WM_E_SHAPE,
               "wm_resize_fwd: geometry H=%d W=%d mid=%dx%d N=%d is outside the fused range (ratio %.2f..%.2f); "
               "use wm_interp_fwd twice", H, W, Hm, Wm, N, RF_RATIO_MIN, RF_RATIO_MAX);
    RFArgs a{};
    a.x = x; a.x_sp = x_sp; a.x_sh = x_sh; a.y = y; a.mask = maskbits; a.mask_wpr = (W + 31) / 32; a.N = N;
    a.ax = RFAxis{W, Wm, (float)W / (float)Wm, (float)Wm / (float)W};
    a.ay = RFAxis{H, Hm, (float)H / (float)Hm, (float)Hm / (float)H};
    a.tiles_x = (W + RF_TW - 1) / RF_TW; a.tiles_y = (H + RF_TH - 1) / RF_TH;
    const size_t smem = rf_smem();
    auto kern = mode == 0 ? resize_fused_fwd_kernel<0> : resize_fused_fwd_kernel<1>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "wm_resize_fwd");
    // persistent over planes: 2 CTAs per SM in total, each tile position walks its share of planes
    const int pos = a.tiles_x * a.tiles_y;
    int gz = (2 * sm_count()) / pos;
    gz = gz < 1 ? 1 : (gz > N ? N : gz);
    kern<<<dim3(a.tiles_x, a.tiles_y, gz), RF_THREADS, smem, (cudaStream_t)stream>>>(a);
    WM_LAUNCH_CHECK("wm_resize_fwd");
    return WM_OK;
}
