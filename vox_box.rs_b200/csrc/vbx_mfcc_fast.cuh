// vbx_mfcc_fast.cuh — MFCC for the common even frame lengths: warps own frames, every stride is a
// compile-time constant.
//
// mfcc_warp_kernel<TIn, TR, MC, R0, R1, R2, R3, FW>: a warp owns FW frames at a time (the (frame, butterfly)
// pairs of a pass are spread over the 32 lanes, so FW = 2 or 4 fills the lanes that a single 25- or 40-butterfly
// pass would leave idle).  The real frame of N = 2·MC samples is packed as MC complex points and transformed by
// a Stockham FFT with radices R0·R1·R2·R3 = MC (radix 8/5/4/3/2 butterflies in registers):
//   pass 0 reads the samples straight from global memory (coalesced), windows them in fp64 and needs no twiddle;
//   later passes run IN PLACE in one shared-memory buffer private to the warp (all inputs of a lane's butterflies
//   are pulled into registers, __syncwarp, then written back) — no CTA barrier after the twiddle table is staged;
// then the packed spectrum is untangled in place, pairwise (k, MC−k), and the band sums / log10 clamp / DCT of
// spectrum.rs:421-439 run in fp64 inside the warp.  Results are identical in structure to mfcc_kernel (the
// any-length CTA kernel); only the summation order inside the FFT differs.
#pragma once

// included from inside vbx_mfcc.cu's anonymous namespace (cxt<>, mk, caddf, csubf, cmulf, mulnegi are defined there)

namespace mfcc_fast {

// shared-memory element index with one pad element per 8: stride-8 (pass 0) writes of 16-byte elements become
// conflict free, contiguous runs stay contiguous inside each 8-block
__device__ __forceinline__ int pidx(int i) { return i + (i >> 3); }

template <int R, typename TR>
__device__ __forceinline__ void bfly(cxt<TR>* v) {
    typedef cxt<TR> C;
    if (R == 2) {
        const C a = v[0], b = v[1];
        v[0] = caddf(a, b);
        v[1] = csubf(a, b);
    } else if (R == 3) {
        const TR s3 = (TR)0.86602540378443864676372317075294;
        const C t1 = caddf(v[1], v[2]);
        const C t2 = mk<TR>(v[0].x - (TR)0.5 * t1.x, v[0].y - (TR)0.5 * t1.y);
        const C d = csubf(v[1], v[2]);
        const C t3 = mulnegi(mk<TR>(s3 * d.x, s3 * d.y));
        v[0] = caddf(v[0], t1);
        v[1] = caddf(t2, t3);
        v[2] = csubf(t2, t3);
    } else if (R == 4) {
        const C a0 = caddf(v[0], v[2]), a1 = csubf(v[0], v[2]);
        const C a2 = caddf(v[1], v[3]), a3 = mulnegi(csubf(v[1], v[3]));
        v[0] = caddf(a0, a2);
        v[1] = caddf(a1, a3);
        v[2] = csubf(a0, a2);
        v[3] = csubf(a1, a3);
    } else if (R == 5) {
        const TR c1 = (TR)0.30901699437494742410229341718282, c2 = (TR)-0.80901699437494742410229341718282;
        const TR s1 = (TR)0.95105651629515357211643933337938, s2 = (TR)0.58778525229247312916870595463907;
        const C t1 = caddf(v[1], v[4]), t2 = caddf(v[2], v[3]);
        const C t3 = csubf(v[1], v[4]), t4 = csubf(v[2], v[3]);
        const C y0 = mk<TR>(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
        const C a1 = mk<TR>(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
        const C a2 = mk<TR>(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
        const C b1 = mulnegi(mk<TR>(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));
        const C b2 = mulnegi(mk<TR>(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
        v[0] = y0;
        v[1] = caddf(a1, b1);
        v[4] = csubf(a1, b1);
        v[2] = caddf(a2, b2);
        v[3] = csubf(a2, b2);
    } else {  // R == 8: three radix-2 stages, W8 = e^{−iπ/4}
        const TR h = (TR)0.70710678118654752440084436210485;
        // stage 1: pairs (k, k+4)
        const C a0 = caddf(v[0], v[4]), a4 = csubf(v[0], v[4]);
        const C a1 = caddf(v[1], v[5]), a5 = csubf(v[1], v[5]);
        const C a2 = caddf(v[2], v[6]), a6 = csubf(v[2], v[6]);
        const C a3 = caddf(v[3], v[7]), a7 = csubf(v[3], v[7]);
        // twiddles on the odd half: a5·W8, a6·W8² = −i·a6, a7·W8³
        const C b5 = mk<TR>(h * (a5.x + a5.y), h * (a5.y - a5.x));
        const C b6 = mulnegi(a6);
        const C b7 = mk<TR>(h * (a7.y - a7.x), -h * (a7.x + a7.y));
        // stage 2 on (a0, a2 | a1, a3) and (a4, b6 | b5, b7)
        const C c0 = caddf(a0, a2), c2 = csubf(a0, a2);
        const C c1 = caddf(a1, a3), c3 = mulnegi(csubf(a1, a3));
        const C c4 = caddf(a4, b6), c6 = csubf(a4, b6);
        const C c5 = caddf(b5, b7), c7 = mulnegi(csubf(b5, b7));
        // stage 3: outputs in natural order X0..X7
        v[0] = caddf(c0, c1);
        v[4] = csubf(c0, c1);
        v[2] = caddf(c2, c3);
        v[6] = csubf(c2, c3);
        v[1] = caddf(c4, c5);
        v[5] = csubf(c4, c5);
        v[3] = caddf(c6, c7);
        v[7] = csubf(c6, c7);
    }
}

// one Stockham pass for FW frames, IN PLACE: every lane first pulls the inputs of all its butterflies into registers
// (NI = ceil(FW·T/32) butterflies of radix R), the warp synchronises, then everybody writes — so a single
// shared-memory buffer per warp suffices.  Radix R, LS = product of the earlier radices, rows MS apart.
// twp = this pass's twiddles laid out [t − 1][k] (t = 1..R−1, k < LS): consecutive lanes read consecutive elements
// (indexing the exp(−2πik/N) table directly walks it with strides that are multiples of 8 elements: 2- to 8-way
// bank conflicts, profiles/r1_mfcc_final_full.txt).
template <typename TR, int MC, int MS, int R, int LS, int FW>
__device__ __forceinline__ void pass_inplace(cxt<TR>* __restrict__ buf, const cxt<TR>* __restrict__ twp, int lane) {
    typedef cxt<TR> C;
    constexpr int T = MC / R;
    constexpr int NI = (FW * T + 31) / 32;
    C v[NI][R];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int item = lane + 32 * i;
        if (item < FW * T) {
            const int q = item / T, j = item - q * T;
            const int k = j % LS;
#pragma unroll
            for (int t = 0; t < R; ++t) {
                v[i][t] = buf[pidx(q * MS + j + t * T)];
                if (t > 0) v[i][t] = cmulf(v[i][t], twp[(t - 1) * LS + k]);
            }
            bfly<R, TR>(v[i]);
        }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int item = lane + 32 * i;
        if (item < FW * T) {
            const int q = item / T, j = item - q * T;
            const int k = j % LS;
            const int o = q * MS + (j - k) * R + k;
#pragma unroll
            for (int u = 0; u < R; ++u) buf[pidx(o + u * LS)] = v[i][u];
        }
    }
    __syncwarp();
}

struct FastParams {
    const void* base;
    const double* win;
    const double2* tw;    // [N] exp(−2πik/N) f64
    const double* wu;
    const double* wd;
    const int* bins;
    const int4* items;    // [num_coeffs + 1] intervals between bin edges {j, first bin, end bin, 0}, longest first
    const double* dct;
    void* out;
    void* energies_out;
    int64_t n_frames, stride, seg_frames, seg_stride;
    int num_coeffs, n_keep, klo, khi;  // stored spectrum indices [klo, khi) ⊆ [0, MC]
    int out_f64;
    int warps_per_cta;
};

template <typename TIn, typename TR, int MC, int R0, int R1, int R2, int R3, int FW>
__global__ void __launch_bounds__(128, 4) mfcc_warp_kernel(const FastParams P) {
    typedef cxt<TR> C;
    static_assert(R0 * R1 * R2 * R3 == MC, "radices must multiply to the transform size");
    constexpr int N = 2 * MC;
    constexpr int MS = MC + 1;  // spectrum row: bins 0..MC
    extern __shared__ __align__(16) unsigned char fast_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int M = P.num_coeffs;
    // twiddles: [MC + 1] of exp(−2πik/N) for the untangle step, then one [R − 1][LS] table per in-place pass
    constexpr int TW1 = (R1 - 1) * R0, TW2 = (R2 - 1) * R0 * R1, TW3 = (R3 - 1) * R0 * R1 * R2;
    constexpr int TWN = MS + TW1 + TW2 + TW3;
    C* s_tw = reinterpret_cast<C*>(fast_smem);                                   // [TWN]
    C* s_tw1 = s_tw + MS;
    C* s_tw2 = s_tw1 + TW1;
    C* s_tw3 = s_tw2 + TW2;
    constexpr int WB = FW * MS + ((FW * MS) >> 3) + 1;                            // padded elements per warp
    C* buf = s_tw + TWN + (size_t)warp * WB;                                     // per warp: FW rows of MS (+ padding)
    const int MD = M + 1;                                                          // padded row of the cosine table (bank spread over k)
    const int ES = 2 * M + 1;                                                      // padded row of the per-frame energies
    double* s_dct = reinterpret_cast<double*>(s_tw + TWN + (size_t)nwarps * WB);  // [n_keep][M + 1] cosine table (CTA wide)
    double2* s_w2 = reinterpret_cast<double2*>(                                    // [N] (rising, falling) weight of bin k, 16-byte aligned
        (reinterpret_cast<uintptr_t>(s_dct + (size_t)P.n_keep * MD) + 15) & ~(uintptr_t)15);
    double* s_e = reinterpret_cast<double*>(s_w2 + N) + (size_t)warp * FW * ES;    // [FW][2M + 1]: rising | falling half-sums, then energies
    auto load_tw = [&](int idx) -> C {
        const double2 w = __ldg(P.tw + idx);
        return mk<TR>((TR)w.x, (TR)w.y);
    };
    for (int i = threadIdx.x; i < MS; i += blockDim.x) s_tw[i] = load_tw(i);
    for (int i = threadIdx.x; i < TW1; i += blockDim.x) s_tw1[i] = load_tw((i % R0) * (i / R0 + 1) * (N / (R0 * R1)));
    for (int i = threadIdx.x; i < TW2; i += blockDim.x) s_tw2[i] = load_tw((i % (R0 * R1)) * (i / (R0 * R1) + 1) * (N / (R0 * R1 * R2)));
    for (int i = threadIdx.x; i < TW3; i += blockDim.x)
        s_tw3[i] = load_tw((i % (R0 * R1 * R2)) * (i / (R0 * R1 * R2) + 1) * (N / (R0 * R1 * R2 * R3)));
    for (int i = threadIdx.x; i < P.n_keep * M; i += blockDim.x) s_dct[(i / M) * MD + (i % M)] = __ldg(P.dct + i);
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_w2[i] = make_double2(__ldg(P.wu + i), __ldg(P.wd + i));
    __syncthreads();

    const int64_t n_groups = (P.n_frames + FW - 1) / FW;
    for (int64_t g = (int64_t)blockIdx.x * nwarps + warp; g < n_groups; g += (int64_t)gridDim.x * nwarps) {
        const int64_t f_first = g * FW;
        const int nf = (int)min((int64_t)FW, P.n_frames - f_first);
        // ---- pass 0: global → registers → radix-R0 butterflies (Ls = 1: no twiddles) → buf -------------------------
        {
            constexpr int T = MC / R0;
#pragma unroll 1
            for (int item = lane; item < FW * T; item += 32) {
                const int q = item / T, j = item - q * T;
                C v[R0];
                if (q < nf) {
                    const int64_t f = f_first + q;
                    const int64_t seg = f / P.seg_frames;
                    const TIn* __restrict__ x = reinterpret_cast<const TIn*>(P.base) + seg * P.seg_stride + (f - seg * P.seg_frames) * P.stride;
#pragma unroll
                    for (int t = 0; t < R0; ++t) {
                        const int i = 2 * (j + t * T);
                        const double2 w = __ldg(reinterpret_cast<const double2*>(P.win + i));
                        v[t] = mk<TR>((TR)((double)vbx_load_sample<TIn>(x + i) * w.x), (TR)((double)vbx_load_sample<TIn>(x + i + 1) * w.y));
                    }
                    bfly<R0, TR>(v);
                } else {
#pragma unroll
                    for (int t = 0; t < R0; ++t) v[t] = mk<TR>((TR)0, (TR)0);  // rows of a ragged last group stay finite
                }
#pragma unroll
                for (int u = 0; u < R0; ++u) buf[pidx(q * MS + j * R0 + u)] = v[u];
            }
            __syncwarp();
        }
        if (R1 > 1) pass_inplace<TR, MC, MS, R1, R0, FW>(buf, s_tw1, lane);
        if (R2 > 1) pass_inplace<TR, MC, MS, R2, R0 * R1, FW>(buf, s_tw2, lane);
        if (R3 > 1) pass_inplace<TR, MC, MS, R3, R0 * R1 * R2, FW>(buf, s_tw3, lane);
        // ---- untangle the packed transform in place, pairwise (k, MC−k): buf[q][k] = (|X_k|², |X_k|), k = 0..MC --------
        {
            constexpr int H = MC / 2 + 1;  // pairs k = 0..MC/2 (k = MC/2 pairs with itself)
#pragma unroll 1
            for (int item = lane; item < FW * H; item += 32) {
                const int q = item / H, k = item - q * H;
                const int zb0 = q * MS;
                const int k2 = MC - k;
                const C za = buf[pidx(zb0 + k)], zb = buf[pidx(zb0 + (k2 == MC ? 0 : k2))];
                // X_k = E − i·T and X_{MC−k} = conj(E) − i·conj(T) with E = (Z_k + conj Z_{MC−k})/2, O = (Z_k − conj Z_{MC−k})/2,
                // T = W_N^k·O  (W_N^{MC−k} = −conj W_N^k, and the pair's even / odd parts are conjugates of each other)
                const C E = mk<TR>((TR)0.5 * (za.x + zb.x), (TR)0.5 * (za.y - zb.y));
                const C O = mk<TR>((TR)0.5 * (za.x - zb.x), (TR)0.5 * (za.y + zb.y));
                const C T = cmulf(s_tw[k], O);
                const TR xr = E.x + T.y, xi = E.y - T.x;      // X_k
                const TR yr = E.x - T.y, yi = -E.y - T.x;     // X_{MC−k}   (k = 0: the Nyquist bin X_MC)
                const TR pwa = xr * xr + xi * xi, pwb = yr * yr + yi * yi;
                const C pa = mk<TR>(pwa, sqrt(pwa)), pb = mk<TR>(pwb, sqrt(pwb));
                buf[pidx(zb0 + k)] = pa;
                buf[pidx(zb0 + k2)] = pb;
            }
            __syncwarp();
        }
        C* dst = buf;
        // ---- band energies (spectrum.rs:421-435): f64 sums, log10, clamp -------------------------------------------------
        // The M + 2 bin edges cut the spectrum into M + 1 intervals; interval j is the rising half of band j (power ×
        // rising weight) and the falling half of band j − 1 (magnitude × the same kind of rising weight — the reference's
        // quirk).  One lane per (interval, frame) reads every bin once for both sums, each in the reference's order; the
        // table lists the intervals longest first so that neighbouring lanes loop about equally long.
#pragma unroll 1
        for (int item = lane; item < FW * (M + 1); item += 32) {
            const int it = item / FW, q = item - it * FW;
            const int4 t = __ldg(P.items + it);
            const int pk0 = q * MS;
            double up = 0., down = 0.;
            for (int k = t.y; k < t.z; ++k) {
                const C sp = dst[pidx(pk0 + ((2 * k > N) ? N - k : k))];
                const double2 w = s_w2[k];
                up = up + (double)sp.x * w.x;
                down = down + (double)sp.y * w.y;
            }
            if (t.x < M) s_e[q * ES + t.x] = up;
            if (t.x >= 1) s_e[q * ES + M + t.x - 1] = down;
        }
        __syncwarp();
#pragma unroll 1
        for (int item = lane; item < FW * M; item += 32) {
            const int q = item / M, w = item - q * M;
            double e = log10(s_e[q * ES + w] + s_e[q * ES + M + w]);
            e = (e > 1.0e-10) ? e : 1.0e-10;  // f64::max(1e-10): NaN → 1e-10
            s_e[q * ES + w] = e;
            if (P.energies_out && q < nf) {
                const size_t o = (size_t)(f_first + q) * M + w;
                if (P.out_f64) reinterpret_cast<double*>(P.energies_out)[o] = e;
                else reinterpret_cast<float*>(P.energies_out)[o] = (float)e;
            }
        }
        __syncwarp();
        // ---- DCT-II ×2, first n_keep rows (spectrum.rs:391-398); four partial sums per row hide the DFMA latency -----------
        const int K = P.n_keep;
#pragma unroll 1
        for (int item = lane; item < FW * K; item += 32) {
            const int q = item / K, k = item - q * K;
            const double* e = s_e + q * ES;
            const double* c = s_dct + k * MD;
            double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
            int m = 0;
            for (; m + 4 <= M; m += 4) {
                a0 = fma(e[m], c[m], a0);
                a1 = fma(e[m + 1], c[m + 1], a1);
                a2 = fma(e[m + 2], c[m + 2], a2);
                a3 = fma(e[m + 3], c[m + 3], a3);
            }
            for (; m < M; ++m) a0 = fma(e[m], c[m], a0);
            if (q < nf) {
                const double v = 2. * ((a0 + a1) + (a2 + a3));
                const size_t o = (size_t)(f_first + q) * K + k;
                if (P.out_f64) reinterpret_cast<double*>(P.out)[o] = v;
                else reinterpret_cast<float*>(P.out)[o] = (float)v;
            }
        }
        __syncwarp();
    }
}

}  // namespace mfcc_fast
