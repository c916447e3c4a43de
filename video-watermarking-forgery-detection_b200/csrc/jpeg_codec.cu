// Real-codec round trip on the device (replaces JpegTest.forward,
// noise_layers/jpeg.py:21-45, which saves every frame through PIL/libjpeg to a temp file).
//
// Huffman coding is lossless, so the pixels libjpeg hands back are an integer function of the
// input bytes; this file computes exactly that function — baseline JPEG as libjpeg(-turbo) does it
// (fixed-point colour conversion, box downsampling with alternating bias, the 13-bit "islow"
// forward/inverse DCT, round-half-away quantisation, "fancy" triangle chroma upsampling) — and
// is bit-exact against Pillow (tests/test_gpu_parity.py, oracle/libjpeg_oracle.py).
//
//   kernel 1 (jc_blocks):   frame tile -> YCbCr bytes in shared memory -> downsample ->
//                           one thread per 8x8 block: fDCT, quantise, dequantise, iDCT ->
//                           decoded Y / Cb / Cr byte planes in a scratch buffer (1.5-3 B/px)
//   kernel 2 (jc_emit):     chroma upsampling across block borders + YCbCr -> RGB -> output
// Integer work on bytes, HBM bound: 12 B/px in, 12 B/px out for float frames, 3 + 3 for bytes.
#include "wm_common.cuh"

namespace wm {
namespace {

constexpr int JC_TW = 256;                 // tile width in full-resolution pixels
constexpr int JC_THREADS = 128;

struct JCTables {
    uint16_t q[2][64];                     // luma / chroma quantisation tables, natural order
    uint32_t magic[2][64];                 // floor(2^32 / (8 q)) + 1: exact division for n < 2^32 / (8 q)
};

struct JCArgs {
    const void* x; int64_t x_sb, x_sc, x_sh;     // input, element strides (W stride 1)
    void* y;                                      // dense [B,3,H,W]
    uint8_t* scratch; int16_t* coef;              // decoded planes; optional quantised coefficients
    int B, H, W, mode;                            // mode: 0 = float [-1,1], 1 = float [0,1], 2 = uint8
    int Hy, Wt, Hc, Wc;                           // padded plane sizes: luma Hy x Wt, chroma Hc x Wc
    int ch, cw;                                   // real (un-padded) chroma size
    int interior_ok;                              // strides / base allow the vector loads of inner tiles
    JCTables t;
};

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

constexpr int C_0_298631336 = 2446, C_0_390180644 = 3196, C_0_541196100 = 4433, C_0_765366865 = 6270;
constexpr int C_0_899976223 = 7373, C_1_175875602 = 9633, C_1_501321110 = 12299, C_1_847759065 = 15137;
constexpr int C_1_961570560 = 16069, C_2_053119869 = 16819, C_2_562915447 = 20995, C_3_072711026 = 25172;

// one 8-point pass of the "islow" forward DCT (jfdctint.c); FIRST: rows, scaled up by 2^PASS1_BITS
template <bool FIRST>
__device__ __forceinline__ void fdct8i(int& d0, int& d1, int& d2, int& d3, int& d4, int& d5, int& d6, int& d7) {
    constexpr int SH = FIRST ? 11 : 15;
    int t0 = d0 + d7, t7 = d0 - d7, t1 = d1 + d6, t6 = d1 - d6;
    int t2 = d2 + d5, t5 = d2 - d5, t3 = d3 + d4, t4 = d3 - d4;
    int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    if (FIRST) { d0 = (t10 + t11) << 2; d4 = (t10 - t11) << 2; }
    else       { d0 = descale(t10 + t11, 2); d4 = descale(t10 - t11, 2); }
    int z1 = (t12 + t13) * C_0_541196100;
    d2 = descale(z1 + t13 * C_0_765366865, SH);
    d6 = descale(z1 - t12 * C_1_847759065, SH);
    z1 = t4 + t7; int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
    int z5 = (z3 + z4) * C_1_175875602;
    t4 *= C_0_298631336; t5 *= C_2_053119869; t6 *= C_3_072711026; t7 *= C_1_501321110;
    z1 *= -C_0_899976223; z2 *= -C_2_562915447;
    z3 = z3 * -C_1_961570560 + z5; z4 = z4 * -C_0_390180644 + z5;
    d7 = descale(t4 + z1 + z3, SH); d5 = descale(t5 + z2 + z4, SH);
    d3 = descale(t6 + z2 + z3, SH); d1 = descale(t7 + z1 + z4, SH);
}

// one 8-point pass of the "islow" inverse DCT (jidctint.c); FIRST: columns
template <bool FIRST>
__device__ __forceinline__ void idct8i(int& c0, int& c1, int& c2, int& c3, int& c4, int& c5, int& c6, int& c7) {
    constexpr int SH = FIRST ? 11 : 18;
    int z1 = (c2 + c6) * C_0_541196100;
    int t2 = z1 - c6 * C_1_847759065, t3 = z1 + c2 * C_0_765366865;
    int t0 = (c0 + c4) << 13, t1 = (c0 - c4) << 13;
    int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    t0 = c7; t1 = c5; t2 = c3; t3 = c1;
    z1 = t0 + t3; int z2 = t1 + t2, z3 = t0 + t2, z4 = t1 + t3;
    int z5 = (z3 + z4) * C_1_175875602;
    t0 *= C_0_298631336; t1 *= C_2_053119869; t2 *= C_3_072711026; t3 *= C_1_501321110;
    z1 *= -C_0_899976223; z2 *= -C_2_562915447;
    z3 = z3 * -C_1_961570560 + z5; z4 = z4 * -C_0_390180644 + z5;
    t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
    c0 = descale(t10 + t3, SH); c7 = descale(t10 - t3, SH);
    c1 = descale(t11 + t2, SH); c6 = descale(t11 - t2, SH);
    c2 = descale(t12 + t1, SH); c5 = descale(t12 - t1, SH);
    c3 = descale(t13 + t0, SH); c4 = descale(t13 - t0, SH);
}

// jdmaster.c prepare_range_limit_table as jidctint.c indexes it (centre folded in, index & 1023)
__device__ __forceinline__ int idct_range_limit(int x) {
    int t = x & 1023;
    return t < 128 ? t + 128 : (t < 512 ? 255 : (t < 896 ? 0 : t - 896));
}

// the frame's byte at (b, c, y, x): JpegTest's own fp32 steps for mode 0
// ((clamp(x,-1,1) + 1) / 2 * 255 truncated, noise_layers/jpeg.py:28), rint(clamp01 * 255) for mode 1
__device__ __forceinline__ int to_byte(float v, int mode) {
    if (mode == 0) {
        v = fminf(fmaxf(v, -1.f), 1.f);
        v = __fmul_rn(__fmul_rn(__fadd_rn(v, 1.f), 0.5f), 255.f);
        return static_cast<int>(v);
    }
    return __float2int_rn(__fmul_rn(fminf(fmaxf(v, 0.f), 1.f), 255.f));
}
__device__ __forceinline__ int fetch_byte(const JCArgs& a, int64_t off) {
    if (a.mode == 2) return static_cast<const uint8_t*>(a.x)[off];
    return to_byte(__ldg(static_cast<const float*>(a.x) + off), a.mode);
}

constexpr int FIXC(double x) { return static_cast<int>(x * 65536.0 + 0.5); }

// jccolor.c rgb_ycc_convert
__device__ __forceinline__ void rgb2ycc(int r, int g, int b, int& y, int& cb, int& cr) {
    constexpr int HALF = 1 << 15, OFF = 128 << 16;
    y  = (FIXC(0.29900) * r + FIXC(0.58700) * g + FIXC(0.11400) * b + HALF) >> 16;
    cb = (-FIXC(0.16874) * r - FIXC(0.33126) * g + FIXC(0.50000) * b + OFF + HALF - 1) >> 16;
    cr = (FIXC(0.50000) * r - FIXC(0.41869) * g - FIXC(0.08131) * b + OFF + HALF - 1) >> 16;
}

// kernel 1: tile of one MCU row (8*VS full-resolution rows x 256 columns)
template <int HS, int VS>
__global__ void __launch_bounds__(JC_THREADS, 5) jc_blocks_kernel(const __grid_constant__ JCArgs a) {
    constexpr int TR = 8 * VS, TW = JC_TW, CW = TW / HS;
    constexpr bool SUB = (HS * VS) > 1;
    __shared__ __align__(16) uint8_t sy[TR][TW];
    __shared__ __align__(16) uint8_t sc[2][TR][TW];             // full-resolution chroma
    __shared__ __align__(16) uint8_t dc[2][8][SUB ? CW : 16];   // downsampled chroma
    __shared__ uint16_t sq[2][64];
    __shared__ uint32_t sm[2][64];

    const int tid = threadIdx.x;
    const int col0 = blockIdx.x * TW, row0 = blockIdx.y * TR, b = blockIdx.z;
    sq[tid >> 6][tid & 63] = a.t.q[tid >> 6][tid & 63];
    sm[tid >> 6][tid & 63] = a.t.magic[tid >> 6][tid & 63];

    // ---- load + colour conversion; edges replicate (jcprepct.c / jcsample.c expand_*_edge) ----
    const int64_t base = (int64_t)b * a.x_sb;
    if (a.interior_ok && col0 + TW <= a.W && row0 + TR <= a.H) {
        // tile inside the frame: four pixels per thread and step, 128-bit (float) / 32-bit (byte) loads
#pragma unroll 4
        for (int i = tid; i < TR * (TW / 4); i += JC_THREADS) {
            const int r = i / (TW / 4), c = (i % (TW / 4)) * 4;
            const int64_t o = base + (int64_t)(row0 + r) * a.x_sh + col0 + c;
            int px[3][4];
            if (a.mode == 2) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(a.x) + o + ch * a.x_sc));
#pragma unroll
                    for (int k = 0; k < 4; ++k) px[ch][k] = (w >> (8 * k)) & 255u;
                }
            } else {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    float4 f = ldg128_stream(static_cast<const float*>(a.x) + o + ch * a.x_sc);
                    px[ch][0] = to_byte(f.x, a.mode); px[ch][1] = to_byte(f.y, a.mode);
                    px[ch][2] = to_byte(f.z, a.mode); px[ch][3] = to_byte(f.w, a.mode);
                }
            }
            uint32_t wy = 0, wb = 0, wr = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int Y, Cb, Cr;
                rgb2ycc(px[0][k], px[1][k], px[2][k], Y, Cb, Cr);
                wy |= (uint32_t)Y << (8 * k); wb |= (uint32_t)Cb << (8 * k); wr |= (uint32_t)Cr << (8 * k);
            }
            *reinterpret_cast<uint32_t*>(&sy[r][c]) = wy;
            *reinterpret_cast<uint32_t*>(&sc[0][r][c]) = wb;
            *reinterpret_cast<uint32_t*>(&sc[1][r][c]) = wr;
        }
    } else {
#pragma unroll 2
        for (int i = tid; i < TR * TW; i += JC_THREADS) {
            int r = i / TW, c = i % TW;
            int gy = row0 + r, fx = min(col0 + c, a.W - 1);
            int fyl = min(gy, a.H - 1);
            // chroma rows past the image repeat the last DOWNSAMPLED row, not the last pixel row
            int fyc = (VS == 2) ? min(2 * min(gy >> 1, a.ch - 1) + (gy & 1), a.H - 1) : fyl;
            int64_t o = base + (int64_t)fyl * a.x_sh + fx;
            int R = fetch_byte(a, o), G = fetch_byte(a, o + a.x_sc), B = fetch_byte(a, o + 2 * a.x_sc);
            int Y, Cb, Cr;
            rgb2ycc(R, G, B, Y, Cb, Cr);
            sy[r][c] = Y;
            if (fyc != fyl) {
                o = base + (int64_t)fyc * a.x_sh + fx;
                R = fetch_byte(a, o); G = fetch_byte(a, o + a.x_sc); B = fetch_byte(a, o + 2 * a.x_sc);
                int Y2;
                rgb2ycc(R, G, B, Y2, Cb, Cr);
            }
            sc[0][r][c] = Cb;
            sc[1][r][c] = Cr;
        }
    }
    __syncthreads();
    // ---- jcsample.c h2v2_downsample (bias 1,2,1,2,...) / h2v1_downsample (bias 0,1,0,1,...) ----
    if (SUB) {
        for (int i = tid; i < 2 * 8 * CW; i += JC_THREADS) {
            int p = i / (8 * CW), r = (i / CW) % 8, c = i % CW;
            int v;
            if (VS == 2)
                v = (sc[p][2 * r][2 * c] + sc[p][2 * r][2 * c + 1] + sc[p][2 * r + 1][2 * c] + sc[p][2 * r + 1][2 * c + 1] +
                     1 + (c & 1)) >> 2;
            else
                v = (sc[p][r][2 * c] + sc[p][r][2 * c + 1] + (c & 1)) >> 1;
            dc[p][r][c] = v;
        }
        __syncthreads();
    }
    // ---- one thread per 8x8 block: fDCT -> quantise -> dequantise -> iDCT, in place ----
    constexpr int NBY = VS * (TW / 8), NBC = CW / 8;
    if (tid < NBY + 2 * NBC) {
        uint8_t* blk; int stride, tsel, gcol, grow, pw, ph; int64_t plane_off;
        if (tid < NBY) {
            int br = tid / (TW / 8), bc = tid % (TW / 8);
            blk = &sy[br * 8][bc * 8]; stride = TW; tsel = 0;
            gcol = col0 + bc * 8; grow = row0 + br * 8; pw = a.Wt; ph = a.Hy;
            plane_off = (int64_t)b * a.Hy * a.Wt;
        } else {
            int k = tid - NBY, p = k / NBC, bc = k % NBC;
            if (SUB) { blk = &dc[p][0][bc * 8]; stride = CW; }
            else     { blk = &sc[p][0][bc * 8]; stride = TW; }
            tsel = 1;
            gcol = col0 / HS + bc * 8; grow = blockIdx.y * 8; pw = a.Wc; ph = a.Hc;
            plane_off = (int64_t)a.B * a.Hy * a.Wt + ((int64_t)p * a.B + b) * a.Hc * a.Wc;
        }
        // blocks that start past the last MCU column are never looked at
        const int lim = tsel ? ((a.cw + 7) & ~7) : (((a.W + 8 * HS - 1) / (8 * HS)) * 8 * HS);
        if (gcol < lim) {
            int v[8][8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                uint2 w = *reinterpret_cast<const uint2*>(blk + r * stride);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    v[r][c]     = (int)((w.x >> (8 * c)) & 255u) - 128;
                    v[r][c + 4] = (int)((w.y >> (8 * c)) & 255u) - 128;
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) fdct8i<true>(v[r][0], v[r][1], v[r][2], v[r][3], v[r][4], v[r][5], v[r][6], v[r][7]);
#pragma unroll
            for (int c = 0; c < 8; ++c) fdct8i<false>(v[0][c], v[1][c], v[2][c], v[3][c], v[4][c], v[5][c], v[6][c], v[7][c]);
            // jcdctmgr.c quantize(): (|c| + 4q) / 8q, sign restored; then jidctint.c's DEQUANTIZE
            int16_t* cf = a.coef ? a.coef + plane_off + (int64_t)grow * pw + gcol : nullptr;
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    int q = sq[tsel][r * 8 + c];
                    int n = abs(v[r][c]) + 4 * q;
                    int m = (int)__umulhi((unsigned)n, sm[tsel][r * 8 + c]);
                    m = v[r][c] < 0 ? -m : m;
                    if (cf) cf[(int64_t)r * pw + c] = (int16_t)m;
                    v[r][c] = m * q;
                }
#pragma unroll
            for (int c = 0; c < 8; ++c) idct8i<true>(v[0][c], v[1][c], v[2][c], v[3][c], v[4][c], v[5][c], v[6][c], v[7][c]);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                idct8i<false>(v[r][0], v[r][1], v[r][2], v[r][3], v[r][4], v[r][5], v[r][6], v[r][7]);
                uint2 w = make_uint2(0u, 0u);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    w.x |= (unsigned)idct_range_limit(v[r][c]) << (8 * c);
                    w.y |= (unsigned)idct_range_limit(v[r][c + 4]) << (8 * c);
                }
                *reinterpret_cast<uint2*>(blk + r * stride) = w;
            }
        }
        (void)ph;
    }
    __syncthreads();
    // ---- decoded planes -> scratch (16-byte stores) ----
    uint8_t* yp = a.scratch + (int64_t)b * a.Hy * a.Wt + (int64_t)row0 * a.Wt + col0;
    for (int i = tid; i < TR * (TW / 16); i += JC_THREADS) {
        int r = i / (TW / 16), c = i % (TW / 16);
        *reinterpret_cast<uint4*>(yp + (int64_t)r * a.Wt + c * 16) = *reinterpret_cast<const uint4*>(&sy[r][c * 16]);
    }
    for (int i = tid; i < 2 * 8 * (CW / 16); i += JC_THREADS) {
        int p = i / (8 * (CW / 16)), r = (i / (CW / 16)) % 8, c = i % (CW / 16);
        uint8_t* cp = a.scratch + (int64_t)a.B * a.Hy * a.Wt + ((int64_t)p * a.B + b) * a.Hc * a.Wc +
                      (int64_t)(blockIdx.y * 8 + r) * a.Wc + col0 / HS + c * 16;
        const uint8_t* s = SUB ? &dc[p][r][c * 16] : &sc[p][r][c * 16];
        *reinterpret_cast<uint4*>(cp) = *reinterpret_cast<const uint4*>(s);
    }
}

// jdcolor.c ycc_rgb_convert (tables folded into arithmetic shifts)
__device__ __forceinline__ void ycc2rgb(int y, int cb, int cr, int& r, int& g, int& b) {
    constexpr int HALF = 1 << 15;
    cb -= 128; cr -= 128;
    r = y + ((FIXC(1.40200) * cr + HALF) >> 16);
    g = y + ((-FIXC(0.34414) * cb + HALF - FIXC(0.71414) * cr) >> 16);
    b = y + ((FIXC(1.77200) * cb + HALF) >> 16);
    r = min(max(r, 0), 255); g = min(max(g, 0), 255); b = min(max(b, 0), 255);
}

// kernel 2: four output pixels per thread.  The fancy upsamplers' first/last-column and
// first/last-row special cases (jdsample.c, jdmainct.c context rows) equal the general formula
// with the neighbour index clamped, which is what is evaluated here.
template <int HS, int VS>
__global__ void __launch_bounds__(256) jc_emit_kernel(const __grid_constant__ JCArgs a) {
    __shared__ float lut[256];
    {
        // ToTensor's u/255 and Normalize's (t - 0.5) / 0.5 in fp32 (noise_layers/jpeg.py:38-43)
        float t = __fdiv_rn((float)threadIdx.x, 255.f);
        lut[threadIdx.x] = a.mode == 0 ? __fdiv_rn(__fsub_rn(t, 0.5f), 0.5f) : t;
    }
    __syncthreads();
    const int x0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int b = blockIdx.z;
    if (x0 >= a.W || y >= a.H) return;
    const uint8_t* yp = a.scratch + (int64_t)b * a.Hy * a.Wt + (int64_t)y * a.Wt + x0;
    const uint32_t yw = *reinterpret_cast<const uint32_t*>(yp);
    int cbv[4], crv[4];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const uint8_t* cp = a.scratch + (int64_t)a.B * a.Hy * a.Wt + ((int64_t)p * a.B + b) * a.Hc * a.Wc;
        int* o = p ? crv : cbv;
        if (HS == 1) {
            uint32_t w = *reinterpret_cast<const uint32_t*>(cp + (int64_t)y * a.Wc + x0);
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = (w >> (8 * i)) & 255u;
        } else {
            const int j0 = x0 >> 1;
            int col[4];                                   // column sums at j0-1 .. j0+2
            if (VS == 2) {
                const int cy = y >> 1;
                const int nb = (y & 1) ? min(cy + 1, a.ch - 1) : max(cy - 1, 0);
                const uint8_t* r0 = cp + (int64_t)cy * a.Wc;
                const uint8_t* r1 = cp + (int64_t)nb * a.Wc;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    int j = min(max(j0 - 1 + i, 0), a.cw - 1);
                    col[i] = 3 * r0[j] + r1[j];
                }
                o[0] = (3 * col[1] + col[0] + 8) >> 4;
                o[1] = (3 * col[1] + col[2] + 7) >> 4;
                o[2] = (3 * col[2] + col[1] + 8) >> 4;
                o[3] = (3 * col[2] + col[3] + 7) >> 4;
            } else {
                const uint8_t* r0 = cp + (int64_t)y * a.Wc;
#pragma unroll
                for (int i = 0; i < 4; ++i) col[i] = r0[min(max(j0 - 1 + i, 0), a.cw - 1)];
                o[0] = (3 * col[1] + col[0] + 1) >> 2;
                o[1] = (3 * col[1] + col[2] + 2) >> 2;
                o[2] = (3 * col[2] + col[1] + 1) >> 2;
                o[3] = (3 * col[2] + col[3] + 2) >> 2;
            }
        }
    }
    int rgb[3][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ycc2rgb((yw >> (8 * i)) & 255u, cbv[i], crv[i], rgb[0][i], rgb[1][i], rgb[2][i]);
    const int64_t hw = (int64_t)a.H * a.W;
    const int64_t o = (int64_t)b * 3 * hw + (int64_t)y * a.W + x0;
    const bool full = (x0 + 4 <= a.W) && (a.W % 4 == 0);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (a.mode == 2) {
            uint8_t* d = static_cast<uint8_t*>(a.y) + o + c * hw;
            if (full) *reinterpret_cast<uint32_t*>(d) = rgb[c][0] | (rgb[c][1] << 8) | (rgb[c][2] << 16) | (rgb[c][3] << 24);
            else for (int i = 0; i < 4 && x0 + i < a.W; ++i) d[i] = rgb[c][i];
        } else {
            float* d = static_cast<float*>(a.y) + o + c * hw;
            if (full) *reinterpret_cast<float4*>(d) = make_float4(lut[rgb[c][0]], lut[rgb[c][1]], lut[rgb[c][2]], lut[rgb[c][3]]);
            else for (int i = 0; i < 4 && x0 + i < a.W; ++i) d[i] = lut[rgb[c][i]];
        }
    }
}

struct JCGeom { int hs, vs, Hy, Wt, Hc, Wc, ch, cw, mcu_rows; };

bool jc_geom(int H, int W, int subsampling, JCGeom& g) {
    if (subsampling == 0) { g.hs = 1; g.vs = 1; }
    else if (subsampling == 1) { g.hs = 2; g.vs = 1; }
    else if (subsampling == 2) { g.hs = 2; g.vs = 2; }
    else return false;
    g.mcu_rows = (H + 8 * g.vs - 1) / (8 * g.vs);
    g.Hy = g.mcu_rows * 8 * g.vs;
    g.Wt = ((W + JC_TW - 1) / JC_TW) * JC_TW;
    g.Hc = g.mcu_rows * 8;
    g.Wc = g.Wt / g.hs;
    g.ch = (H + g.vs - 1) / g.vs;
    g.cw = (W + g.hs - 1) / g.hs;
    return true;
}

// jcparam.c jpeg_quality_scaling + jpeg_add_quant_table(force_baseline)
void jc_tables(int quality, JCTables& t) {
    static const int luma[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55,
                                 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                                 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                                 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
    static const int chroma4[4][4] = {{17, 18, 24, 47}, {18, 21, 26, 66}, {24, 26, 56, 99}, {47, 66, 99, 99}};
    int q = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
    int scale = q < 50 ? 5000 / q : 200 - 2 * q;
    for (int i = 0; i < 64; ++i) {
        int r = i / 8, c = i % 8;
        int std2[2] = {luma[i], (r < 4 && c < 4) ? chroma4[r][c] : 99};
        for (int p = 0; p < 2; ++p) {
            int v = (std2[p] * scale + 50) / 100;
            v = v < 1 ? 1 : (v > 255 ? 255 : v);
            t.q[p][i] = (uint16_t)v;
            t.magic[p][i] = (uint32_t)((1ull << 32) / (uint64_t)(8 * v)) + 1u;
        }
    }
}

}  // namespace
}  // namespace wm

using namespace wm;

extern "C" int64_t wm_jpegcodec_scratch_bytes(int B, int H, int W, int subsampling) {
    JCGeom g;
    if (B <= 0 || H <= 0 || W <= 0 || !jc_geom(H, W, subsampling, g)) return 0;
    return (int64_t)B * ((int64_t)g.Hy * g.Wt + 2ll * g.Hc * g.Wc);
}

extern "C" int wm_jpegcodec(const void* x, int64_t x_sb, int64_t x_sc, int64_t x_sh, void* y, int B, int H, int W,
                            int quality, int subsampling, int mode, uint8_t* scratch, int16_t* coef, void* stream) {
    if (B <= 0 || H <= 0 || W <= 0) return WM_OK;      // empty work
    WM_REQUIRE(x && y && scratch, WM_E_NULL, "wm_jpegcodec: null pointer");
    WM_REQUIRE(mode >= 0 && mode <= 2, WM_E_ARG, "wm_jpegcodec: mode must be 0 ([-1,1] float), 1 ([0,1] float) or 2 (uint8)");
    JCGeom g;
    WM_REQUIRE(jc_geom(H, W, subsampling, g), WM_E_ARG, "wm_jpegcodec: subsampling must be 0 (4:4:4), 1 (4:2:2) or 2 (4:2:0)");
    WM_REQUIRE(quality >= 1 && quality <= 100, WM_E_ARG, "wm_jpegcodec: quality %d outside 1..100", quality);
    WM_REQUIRE(B <= 65535 && g.mcu_rows <= 65535 && (H + 3) / 4 <= 65535, WM_E_SHAPE, "wm_jpegcodec: batch / height too large for one launch");
    WM_REQUIRE(aligned(scratch, 16) && aligned(y, 16), WM_E_ALIGN, "wm_jpegcodec: scratch and y must be 16-byte aligned");
    JCArgs a{};
    a.x = x; a.x_sb = x_sb; a.x_sc = x_sc; a.x_sh = x_sh; a.y = y; a.scratch = scratch; a.coef = coef;
    a.B = B; a.H = H; a.W = W; a.mode = mode;
    a.Hy = g.Hy; a.Wt = g.Wt; a.Hc = g.Hc; a.Wc = g.Wc; a.ch = g.ch; a.cw = g.cw;
    {
        const size_t el = mode == 2 ? 1 : 4, need = mode == 2 ? 4 : 16;
        a.interior_ok = aligned(x, need) && (x_sb * el) % need == 0 && (x_sc * el) % need == 0 && (x_sh * el) % need == 0;
    }
    jc_tables(quality, a.t);
    cudaStream_t s = (cudaStream_t)stream;
    dim3 g1(g.Wt / JC_TW, g.mcu_rows, B), g2((W + 255) / 256, (H + 3) / 4, B);
    if (subsampling == 2)      { jc_blocks_kernel<2, 2><<<g1, JC_THREADS, 0, s>>>(a); jc_emit_kernel<2, 2><<<g2, 256, 0, s>>>(a); }
    else if (subsampling == 1) { jc_blocks_kernel<2, 1><<<g1, JC_THREADS, 0, s>>>(a); jc_emit_kernel<2, 1><<<g2, 256, 0, s>>>(a); }
    else                       { jc_blocks_kernel<1, 1><<<g1, JC_THREADS, 0, s>>>(a); jc_emit_kernel<1, 1><<<g2, 256, 0, s>>>(a); }
    WM_LAUNCH_CHECK("wm_jpegcodec");
    return WM_OK;
}
