// Fused log-mel feature extraction for sm_100a.
//
// Replaces, in ONE kernel and with each audio sample read from HBM once
// (overlapping frames are served from L1/L2), the reference call chain
//   F.pad(zeros 432)            emphases/core.py:357-358
//   audio[:, s:e] chunk slice   emphases/core.py:395-401
//   F.pad(reflect 432)          emphases/data/preprocess/mels.py:32-36
//   torch.stft(1024, hop 160, hann, center=False)      mels.py:39-48
//   sqrt(re^2 + im^2 + 1e-6)    mels.py:51
//   basis @ spectrogram, log(clamp(., 1e-5)), optional (x+10)/10
//                               mels.py:94-109, 57-58
//
// Persistent CTAs of 8 warps, two per SM (so one CTA's mel / store phases overlap
// the other's FFT phase), each working on tiles of 16 consecutive packed rows:
//   phase 1  one warp per frame: 1024-point real FFT as a 512-point complex
//            FFT of the even/odd packed signal -- 3 radix-8 Stockham passes
//            (2 butterflies per lane per pass, twiddles held in registers,
//            two exchanges through bank-padded shared memory), then the
//            real-FFT unpacking done in registers with warp shuffles and the
//            magnitudes written to a [16 frames][513 bins] tile;
//   phase 2  mel projection with lane = frame (a half-warp per mel row).  The
//            CSR basis is expanded once per CTA into a banded dense table: a
//            row's non-zeros (a triangle = one contiguous run of bins) are
//            widened to 4-bin-aligned groups, and the two rows a warp works
//            on together are padded to the same group count, so the inner
//            loop is branch-uniform: one 16-byte weight broadcast + one
//            16-byte magnitude load (row stride 516 floats: conflict-free
//            per quarter-warp) feed 4 FMAs;
//   phase 3  log / clamp and a fully coalesced store of the 16 x n_mels tile.
// All fp32.  Bound: FP32 issue + shared-memory wavefronts (about 25 kFLOP per
// frame against 960 B of HBM traffic), see DESIGN.md.
#include "common.cuh"

namespace emph {

constexpr int kFft = 1024;
constexpr int kHop = 160;
constexpr int kPad = (kFft - kHop) / 2;   // 432, both the zero and reflect pad
constexpr int kBins = kFft / 2 + 1;       // 513
constexpr int kHalf = kFft / 2;           // 512-point complex FFT
constexpr int kWarps = 8;                  // two CTAs per SM: their phases interleave
constexpr int kTile = 16;                 // frames per CTA tile
constexpr int kMaxMels = 128;
constexpr int kMagStride = 516;           // floats; 16-byte aligned rows, 129 = 1 mod 8 quads
constexpr int kMaxQuads = 768;            // banded mel table: float4 groups held in smem
constexpr int kSpan = (kTile - 1) * kHop + kFft;   // samples one interior tile reads: 3424
constexpr int kRound = 256;               // tile descriptors are computed 256 tiles ahead

// Exchange buffer of one warp (float2 slots).  The two exchanges use different
// placements, both chosen so that every 8-byte access of a half-warp touches 16
// distinct banks AND every address is a per-lane base plus a compile-time
// constant:
//   after pass 1, element (butterfly j, output r), logical index 8 j + r, sits at
//     66 r + j                        (written lane = j, read lane % 8 = r)
//   after pass 2, logical index 64 a + 8 r + b sits at
//     b + 8 (a % 4) + 40 (r % 4) + 152 (a / 4) + 304 (r / 4)
//                                      (written lane = 8 (a % 4) + b, read
//                                       lane = 8 (r % 4) + b)
constexpr int kXchg = 608;

struct __align__(16) LogmelSmem {
    float mag[kTile][kMagStride];  // bins 513..515 stay zero (band padding reads them)
    float2 xchg[kWarps][kXchg];
    float hann[kFft];              // 0.5 * hann: the real-FFT unpacking's 1/2 is folded in
    float outs[kTile][kMaxMels + 1];
    float4 mel_quad[kMaxQuads];    // banded dense weights, 4 bins per entry
    int32_t mel_first[kMaxMels];   // first quad of the row in mel_quad
    int32_t mel_bin0[kMaxMels];    // first bin of the row's band (multiple of 4)
    int32_t mel_quads[kMaxMels / 2];   // double-quads per row of the pair (both rows padded to it)
    int32_t mel_fits;              // 0: the basis does not fit the table, use the CSR path
    // audio of one interior tile (16 frames = 3424 samples), fetched with one
    // cp.async.bulk while the previous tile is in its mel / store phases
    __align__(16) unsigned char stage[kSpan * sizeof(float)];
    long long tile_src[kRound];    // per tile of this CTA: element index of the span, or -1
    unsigned long long bar;        // mbarrier the bulk copy completes on
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
// one elected thread: expect `bytes` on the barrier and start the bulk copy
__device__ __forceinline__ void bulk_fetch(
    void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile(
        "{\n\t.reg .b64 state;\n\t"
        "mbarrier.arrive.expect_tx.shared::cta.b64 state, [%0], %1;\n\t}"
        ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}


__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void bfly2(float2& a, float2& b) {
    float2 t = a;
    a = make_float2(t.x + b.x, t.y + b.y);
    b = make_float2(t.x - b.x, t.y - b.y);
}
// sqrt of a strictly positive normal number (x >= 1e-6 here): one MUFU.SQRT
// (relative error <= 2^-23, far inside the 2e-5 log-mel tolerance)
__device__ __forceinline__ float sqrt_pos(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// natural log of a positive normal number (x >= 1e-5 here): MUFU.LG2 * ln 2,
// absolute error ~1e-7 relative to log-mel values of magnitude 1..12
__device__ __forceinline__ float log_pos(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * 0.69314718055994530942f;
}
// multiply by -i
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

// 4-point DFT; outputs land as (X0, X2, X1, X3) in (a, b, c, d)
__device__ __forceinline__ void fft4(float2& a, float2& b, float2& c, float2& d) {
    bfly2(a, c);
    bfly2(b, d);
    d = mul_mi(d);
    bfly2(a, b);
    bfly2(c, d);
}

// 8-point DFT in place; v[] ends up in natural order
__device__ __forceinline__ void fft8(float2 (&v)[8]) {
    const float s = 0.70710678118654752440f;
    bfly2(v[0], v[4]);
    bfly2(v[1], v[5]);
    bfly2(v[2], v[6]);
    bfly2(v[3], v[7]);
    v[5] = make_float2(s * (v[5].x + v[5].y), s * (v[5].y - v[5].x));     // * (s, -s)
    v[6] = mul_mi(v[6]);
    v[7] = make_float2(s * (v[7].y - v[7].x), -s * (v[7].x + v[7].y));    // * (-s, -s)
    fft4(v[0], v[1], v[2], v[3]);   // X0 X4 X2 X6
    fft4(v[4], v[5], v[6], v[7]);   // X1 X5 X3 X7
    float2 t1 = v[1], t3 = v[3], t4 = v[4], t6 = v[6];
    v[1] = t4;      // X1
    v[3] = t6;      // X3
    v[4] = t1;      // X4
    v[6] = t3;      // X6
}

template <typename T>
__device__ __forceinline__ float to_float(T v);
template <>
__device__ __forceinline__ float to_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_float<int16_t>(int16_t v) {
    return (float)v * (1.f / 32768.f);
}

// Sample q of the reflect-padded chunk, through the whole index map
// (SURVEY.md A.2): chunk[j] = P[s + j], P = zeros(432) ++ audio ++ zeros(432)
template <typename T>
__device__ __forceinline__ float chunk_sample(
    const T* __restrict__ audio, int T_len, int s, int L, int q) {
    int j = q - kPad;
    if (j < 0) j = -j;
    if (j >= L) j = 2 * (L - 1) - j;
    int a = s + j - kPad;
    return (a >= 0 && a < T_len) ? to_float<T>(audio[a]) : 0.f;
}

template <typename T>
__global__ void __launch_bounds__(kWarps * 32, 2)
logmel_kernel(
    const T* __restrict__ audio,
    const int64_t* __restrict__ audio_off, const int32_t* __restrict__ audio_len,
    const int32_t* __restrict__ chunk_start, const int32_t* __restrict__ chunk_len,
    const int32_t* __restrict__ row_start,
    const int32_t* __restrict__ row_seq, int total_rows,
    const int32_t* __restrict__ mel_ptr, const int16_t* __restrict__ mel_col,
    const float* __restrict__ mel_val, int n_mels, int normalize,
    float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LogmelSmem& sm = *reinterpret_cast<LogmelSmem*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < kFft; i += blockDim.x) {
        // torch.hann_window(1024) (periodic): 0.5 - 0.5 cos(2 pi n / N), times the
        // 1/2 of the real-FFT unpacking (exact: a power of two)
        sm.hann[i] = 0.25f - 0.25f * cospif(2.f * (float)i / (float)kFft);
    }

    // ---- banded mel table, built once per CTA from the CSR basis ----
    for (int i = tid; i < kTile * (kMagStride - kBins); i += blockDim.x)
        sm.mag[i / (kMagStride - kBins)][kBins + i % (kMagStride - kBins)] = 0.f;
    for (int i = tid; i < kMaxQuads; i += blockDim.x)
        sm.mel_quad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int m = tid; m < n_mels; m += blockDim.x) {
        // band of row m: [first bin rounded down to 4, last bin]; empty row: 0 quads
        const int e0 = mel_ptr[m], e1 = mel_ptr[m + 1];
        int lo = kBins, hi = -1;
        for (int e = e0; e < e1; ++e) {
            const int c = mel_col[e];
            lo = min(lo, c);
            hi = max(hi, c);
        }
        const bool valid = hi >= lo && lo >= 0 && hi < kBins;
        sm.mel_bin0[m] = valid ? (lo & ~3) : 0;
        sm.mel_first[m] = valid ? (hi - (lo & ~3)) / 4 + 1 : (e1 > e0 ? kMaxQuads + 1 : 0);
    }
    __syncthreads();
    if (tid == 0) {
        const int n_pairs = (n_mels + 1) >> 1;
        int total = 0;
        bool fits = true;
        for (int p = 0; p < n_pairs; ++p) {
            const int m0 = 2 * p, m1 = min(2 * p + 1, n_mels - 1);
            // (an even count: the inner loop takes two quads per step)
            const int quads = (max(sm.mel_first[m0], sm.mel_first[m1]) + 1) & ~1;
            if (4 * quads > kMagStride || total + 2 * quads > kMaxQuads) {
                fits = false;
                break;
            }
            sm.mel_quads[p] = quads >> 1;
            // a band padded past the end of the mag row is moved down instead
            sm.mel_bin0[m0] = min(sm.mel_bin0[m0], kMagStride - 4 * quads);
            sm.mel_first[m0] = total;
            if (m1 != m0) {
                sm.mel_bin0[m1] = min(sm.mel_bin0[m1], kMagStride - 4 * quads);
                sm.mel_first[m1] = total + quads;
            }
            total += 2 * quads;
        }
        sm.mel_fits = fits ? 1 : 0;
    }
    __syncthreads();
    const bool banded = sm.mel_fits != 0;
    if (banded) {
        float* const table = reinterpret_cast<float*>(sm.mel_quad);
        for (int m = warp; m < n_mels; m += kWarps) {
            const int base = 4 * sm.mel_first[m] - sm.mel_bin0[m];
            for (int e = mel_ptr[m] + lane; e < mel_ptr[m + 1]; e += 32)
                table[base + mel_col[e]] = mel_val[e];
        }
    }

    // Per-lane twiddles, kept in registers for the whole kernel:
    //   pass 2: exp(-2 pi i * 8 r (lane % 8) / 512)
    //   pass 3: exp(-2 pi i r j / 512) for j = lane and j = 64 - lane (lane 0: 32)
    //   unpack: exp(-2 pi i k / 1024), k the bin of slot q
    float2 tw2[8], tw3[8], tw3b[8], wq[8];
#pragma unroll
    for (int r = 1; r < 8; ++r) {
        float sn, cs;
        sincospif(-2.f * (float)(8 * r * (lane & 7)) / (float)kHalf, &sn, &cs);
        tw2[r] = make_float2(cs, sn);
        sincospif(-2.f * (float)(r * lane) / (float)kHalf, &sn, &cs);
        tw3[r] = make_float2(cs, sn);
        sincospif(-2.f * (float)(r * (lane ? 64 - lane : 32)) / (float)kHalf, &sn, &cs);
        tw3b[r] = make_float2(cs, sn);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        // bin of unpacking slot q: lane + 64 q; lane 0 takes 64 q (q < 4) and 32 + 64 (q - 4)
        const int k = (lane == 0 && q >= 4) ? 64 * q - 224 : lane + 64 * q;
        sincospif(-2.f * (float)k / (float)kFft, &wq[q].y, &wq[q].x);
    }
    __syncthreads();

    float2* const xw = sm.xchg[warp];
    float2* const pw = xw + lane;                               // both stores
    float2* const pr2 = xw + (lane & 7) * 66 + (lane >> 3);     // pass-2 loads
    float2* const pr3 = xw + (lane & 7) + 40 * (lane >> 3);     // pass-3 loads, butterfly j = lane
    // pass 3's second butterfly is j = 64 - lane (lane 0: 32), so that Z[k] and
    // Z[512 - k] of the real-FFT unpacking meet in one lane
    const int j1 = lane ? 64 - lane : 32;
    float2* const pr3b = xw + (j1 & 7) + 40 * ((j1 >> 3) & 3) + 304;
    const float4* const ph = reinterpret_cast<const float4*>(sm.hann) + lane;
    const int n_tiles = (total_rows + kTile - 1) / kTile;

    // Tile descriptors, kRound tiles of this CTA at a time: the element index of
    // the tile's first sample when all 16 rows are interior frames of one chunk
    // (no separator, no zero / reflect padding, span inside the utterance) and
    // the span is 16-byte aligned; -1 sends the tile down the per-frame path.
    auto describe_round = [&](int first_it) {
        const long long t = (long long)blockIdx.x + (long long)(first_it + tid) * gridDim.x;
        long long desc = -1;
        const long long r0 = t * kTile;
        if (tid < kRound && r0 + kTile <= total_rows) {
            const int u = __ldg(row_seq + r0);
            if (u >= 0 && __ldg(row_seq + r0 + kTile - 1) == u) {
                const int frame = (int)r0 - __ldg(row_start + u);
                const int q0 = frame * kHop;
                const int a0 = __ldg(chunk_start + u) + q0 - 2 * kPad;
                const long long index = __ldg(audio_off + u) + a0;
                const bool ok =
                    q0 >= kPad && q0 + kSpan - kPad <= __ldg(chunk_len + u) &&
                    a0 >= 0 && a0 + kSpan <= __ldg(audio_len + u) &&
                    (reinterpret_cast<uintptr_t>(audio + index) & 15) == 0;
                if (ok) desc = index;
            }
        }
        if (tid < kRound) sm.tile_src[tid] = desc;
    };
    if (tid == 0) mbar_init(&sm.bar, 1);
    describe_round(0);
    __syncthreads();
    if (tid == 0 && sm.tile_src[0] >= 0)
        bulk_fetch(sm.stage, audio + sm.tile_src[0], kSpan * sizeof(T), &sm.bar);
    uint32_t parity = 0;
    int it = 0;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int row0 = tile * kTile;
        const bool staged = sm.tile_src[it & (kRound - 1)] >= 0;
        if (staged) {
            mbar_wait(&sm.bar, parity);
            parity ^= 1;
        }

        // ======================= phase 1: FFT + magnitude =======================
#pragma unroll 1
        for (int f = warp; f < kTile; f += kWarps) {
            // ---- load 1024 samples as 512 complex, window, first radix-8 pass ----
            // pass 1: lane handles butterflies j = 2 lane (v0) and 2 lane + 1 (v1);
            // inputs z[j + 64 r], so one 16-byte load brings both butterflies' inputs
            float2 v0[8], v1[8];
            auto load_run = [&](const T* p) {    // 1024 contiguous samples
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int n = 4 * lane + 128 * r;
                    if constexpr (sizeof(T) == 4) {
                        const float4 x = *reinterpret_cast<const float4*>(p + n);
                        v0[r] = make_float2(x.x, x.y);
                        v1[r] = make_float2(x.z, x.w);
                    } else {
                        const short4 x = *reinterpret_cast<const short4*>(p + n);
                        v0[r] = make_float2(to_float<int16_t>(x.x), to_float<int16_t>(x.y));
                        v1[r] = make_float2(to_float<int16_t>(x.z), to_float<int16_t>(x.w));
                    }
                }
            };
            if (staged) {
                load_run(reinterpret_cast<const T*>(sm.stage) + f * kHop);
            } else {
                const int row = row0 + f;
                if (row >= total_rows) break;
                const int u = __ldg(row_seq + row);
                if (u < 0) continue;                    // separator row
                const int frame = row - __ldg(row_start + u);
                const int T_len = __ldg(audio_len + u);
                const int s = __ldg(chunk_start + u);
                const int L = __ldg(chunk_len + u);
                const T* src = audio + __ldg(audio_off + u);
                const int q0 = frame * kHop;            // first sample in reflect-padded coords
                const int a0 = s + q0 - 2 * kPad;       // audio index of sample q0
                const bool interior = (q0 >= kPad) && (q0 + kFft - kPad <= L) &&
                                      (a0 >= 0) && (a0 + kFft <= T_len);
                if (interior) {
                    load_run(src + a0);
                } else {
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const int n = q0 + 4 * lane + 128 * r;
                        v0[r] = make_float2(
                            chunk_sample<T>(src, T_len, s, L, n),
                            chunk_sample<T>(src, T_len, s, L, n + 1));
                        v1[r] = make_float2(
                            chunk_sample<T>(src, T_len, s, L, n + 2),
                            chunk_sample<T>(src, T_len, s, L, n + 3));
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float4 h = ph[32 * r];
                v0[r].x *= h.x; v0[r].y *= h.y;
                v1[r].x *= h.z; v1[r].y *= h.w;
            }
            // pass 1: Ns = 1, no twiddles, out[8 j + r]
            fft8(v0);
            fft8(v1);
#pragma unroll
            for (int r = 0; r < 8; ++r)     // slots 66 r + 2 lane, + 1: one 16-byte store
                *reinterpret_cast<float4*>(pw + lane + 66 * r) =
                    make_float4(v0[r].x, v0[r].y, v1[r].x, v1[r].y);
            __syncwarp();

            // pass 2: Ns = 8; twiddle exp(-2 pi i r (j % 8) / 64)
            {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    v0[r] = pr2[8 * r];
                    v1[r] = pr2[4 + 8 * r];
                }
#pragma unroll
                for (int r = 1; r < 8; ++r) {
                    v0[r] = cmul(v0[r], tw2[r]);
                    v1[r] = cmul(v1[r], tw2[r]);
                }
                fft8(v0);
                fft8(v1);
                __syncwarp();
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    pw[40 * (r & 3) + 304 * (r >> 2)] = v0[r];
                    pw[152 + 40 * (r & 3) + 304 * (r >> 2)] = v1[r];
                }
                __syncwarp();
            }

            // pass 3: Ns = 64; twiddle exp(-2 pi i r j / 512), results stay in
            // registers: v0[r] = Z[lane + 64 r], v1[r] = Z[j1 + 64 r]
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                v0[r] = pr3[8 * (r & 3) + 152 * (r >> 2)];
                v1[r] = pr3b[8 * (r & 3) + 152 * (r >> 2)];
            }
            __syncwarp();                      // xw is free for the next frame
#pragma unroll
            for (int r = 1; r < 8; ++r) {
                v0[r] = cmul(v0[r], tw3[r]);
                v1[r] = cmul(v1[r], tw3b[r]);
            }
            fft8(v0);
            fft8(v1);

            // ---- real-FFT unpacking in registers ----
            // With e = (Z[k] + conj Z[512-k]) / 2, o = -i (Z[k] - conj Z[512-k]) / 2
            // (the 1/2 is already in the window) and w = exp(-2 pi i k / 1024):
            // X[k] = e + w o,  X[512-k] = conj(e - w o),
            // so one (e, w o) serves both bins of a pair.  (The magnitudes are taken
            // from the complex sums, not as |e|^2 + |o|^2 +- 2 Re(e conj(w o)):
            // bins k and 512-k of speech differ by up to 60 dB and subtracting
            // powers would wipe out the weak one.)
            // Slot q of a lane != 0 pairs k = lane + 64 q (v0[q]) with
            // 512 - k = (64 - lane) + 64 (7 - q) (v1[7 - q]): no shuffles.  Lane 0
            // holds the self-paired residues 0 and 32: slots 0..3 pair v0[q] with
            // v0[(8 - q) % 8] (k = 64 q), slots 4..7 pair v1[q - 4] with v1[11 - q]
            // (k = 32 + 64 (q - 4)); v0[4] = Z[256] is its own partner (below).
            float* const mag = sm.mag[f];
            const bool lane0 = lane == 0;
            float* const mag_lo = mag + lane;                       // slots 0..3: mag_lo[64 q]
            float* const mag_hi = mag + (lane0 ? -224 : lane);      // slots 4..7: mag_hi[64 q]
            float* const mir_lo = mag - lane;                       // mir_lo[512 - 64 q]
            float* const mir_hi = mag - (lane0 ? -224 : lane);      // mir_hi[512 - 64 q]
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float2 a = v0[q], pz = v1[7 - q];
                if (q < 4) {
                    const float2 alt = v0[(8 - q) & 7];
                    pz.x = lane0 ? alt.x : pz.x;
                    pz.y = lane0 ? alt.y : pz.y;
                } else {
                    const float2 alt_a = v1[q - 4], alt = v1[11 - q];
                    a.x = lane0 ? alt_a.x : a.x;
                    a.y = lane0 ? alt_a.y : a.y;
                    pz.x = lane0 ? alt.x : pz.x;
                    pz.y = lane0 ? alt.y : pz.y;
                }
                const float2 e = make_float2(a.x + pz.x, a.y - pz.y);
                const float2 o = make_float2(a.y + pz.y, pz.x - a.x);       // -i * (a - conj pz)
                const float2 wo = cmul(wq[q], o);
                const float xr = e.x + wo.x, xi = e.y + wo.y;
                const float yr = e.x - wo.x, yi = e.y - wo.y;
                (q < 4 ? mag_lo : mag_hi)[64 * q] = sqrt_pos(fmaf(xr, xr, fmaf(xi, xi, 1e-6f)));
                (q < 4 ? mir_lo : mir_hi)[kHalf - 64 * q] = sqrt_pos(fmaf(yr, yr, fmaf(yi, yi, 1e-6f)));
            }
            // k = 256 pairs with itself: X[256] = conj(Z[256]); Z[256] = Z_8 of lane 0
            // (Z is halved by the window: |X|^2 = 4 |Z|^2)
            if (lane0) mag[kHalf / 2] = sqrt_pos(fmaf(4.f * v0[4].x, v0[4].x, fmaf(4.f * v0[4].y, v0[4].y, 1e-6f)));
        }
        __syncthreads();

        // every warp is done with the staged audio: fetch the next tile's span
        // (it lands while this tile is in its mel / store phases)
        if (((it + 1) & (kRound - 1)) == 0) {
            describe_round(it + 1);
            __syncthreads();
        }
        if (tid == 0 && tile + (int)gridDim.x < n_tiles) {
            const long long next = sm.tile_src[(it + 1) & (kRound - 1)];
            if (next >= 0) bulk_fetch(sm.stage, audio + next, kSpan * sizeof(T), &sm.bar);
        }

        // ============ phase 2: banded mel projection, half-warp lane = frame ============
        {
            const int f = lane & (kTile - 1);          // frame of this lane
            const int h = lane >> 4;                   // which row of the pair
            const int row = row0 + f;
            const bool live = staged || (row < total_rows && __ldg(row_seq + row) >= 0);
            const int n_pairs = (n_mels + 1) >> 1;
            // pairs are dealt to warps in a snake so wide (high-frequency) and
            // narrow (low-frequency) filters balance
            for (int j = 0; j * kWarps < n_pairs; ++j) {
                const int p = j * kWarps + ((j & 1) ? kWarps - 1 - warp : warp);
                if (p >= n_pairs) continue;
                const int m = min(2 * p + h, n_mels - 1);
                float acc;
                if (banded) {
                    const int steps = sm.mel_quads[p];
                    const float4* wt = sm.mel_quad + sm.mel_first[m];
                    const float4* xq = reinterpret_cast<const float4*>(sm.mag[f]) +
                                       (sm.mel_bin0[m] >> 2);
                    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll 1
                    for (const float4* const end = wt + 2 * steps; wt != end; wt += 2, xq += 2) {
                        const float4 w0 = wt[0], x0 = xq[0];
                        const float4 w1 = wt[1], x1 = xq[1];
                        acc0 = fmaf(w0.x, x0.x, acc0);
                        acc1 = fmaf(w0.y, x0.y, acc1);
                        acc2 = fmaf(w0.z, x0.z, acc2);
                        acc3 = fmaf(w0.w, x0.w, acc3);
                        acc0 = fmaf(w1.x, x1.x, acc0);
                        acc1 = fmaf(w1.y, x1.y, acc1);
                        acc2 = fmaf(w1.z, x1.z, acc2);
                        acc3 = fmaf(w1.w, x1.w, acc3);
                    }
                    acc = (acc0 + acc1) + (acc2 + acc3);
                } else {
                    // any other basis: CSR entries straight from global memory
                    acc = 0.f;
                    for (int e = __ldg(mel_ptr + m); e < __ldg(mel_ptr + m + 1); ++e)
                        acc = fmaf(__ldg(mel_val + e), sm.mag[f][__ldg(mel_col + e)], acc);
                }
                float v = log_pos(fmaxf(acc, 1e-5f));
                if (normalize) v = (v + 10.f) / 10.f;
                sm.outs[f][m] = live ? v : 0.f;
            }
        }
        __syncthreads();

        // ========================= phase 3: coalesced store =========================
        {
            const int rows = min(kTile, total_rows - row0);
            float* dst = out + (size_t)row0 * n_mels;
            for (int f = warp; f < rows; f += kWarps)
                for (int c = lane; c < n_mels; c += 32)
                    dst[f * n_mels + c] = sm.outs[f][c];
        }
        // the next tile's phase 1 only touches mag / xchg; outs is rewritten
        // after the next __syncthreads
    }
}

template <typename T>
int launch_logmel(
    const T* audio,
    const int64_t* audio_off, const int32_t* audio_len,
    const int32_t* chunk_start, const int32_t* chunk_len,
    const int32_t* row_start, int32_t n_seq,
    const int32_t* row_seq, int32_t total_rows,
    const int32_t* mel_ptr, const int16_t* mel_col, const float* mel_val,
    int32_t n_mels, int32_t normalize, float* out, void* stream) {
    EMPH_REQUIRE(n_seq >= 0 && total_rows >= 0, "emph_logmel: negative size");
    EMPH_REQUIRE(n_mels > 0 && n_mels <= kMaxMels, "emph_logmel: n_mels %d out of range", n_mels);
    if (total_rows == 0) return EMPH_OK;
    const size_t smem = sizeof(LogmelSmem);
    int s = check_cuda(
        cudaFuncSetAttribute(
            logmel_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
        "logmel smem attribute");
    if (s != EMPH_OK) return s;
    const int n_tiles = (total_rows + kTile - 1) / kTile;
    const int grid = n_tiles < 2 * sm_count() ? n_tiles : 2 * sm_count();
    logmel_kernel<T><<<grid, kWarps * 32, smem, (cudaStream_t)stream>>>(
        audio, audio_off, audio_len, chunk_start, chunk_len, row_start,
        row_seq, total_rows, mel_ptr, mel_col, mel_val, n_mels, normalize, out);
    EMPH_CHECK_LAUNCH("emph_logmel");
    return EMPH_OK;
}

}  // namespace emph

extern "C" {

int emph_logmel_f32(
    const float* audio,
    const int64_t* audio_off, const int32_t* audio_len,
    const int32_t* chunk_start, const int32_t* chunk_len,
    const int32_t* row_start, int32_t n_seq,
    const int32_t* row_seq, int32_t total_rows,
    const int32_t* mel_ptr, const int16_t* mel_col, const float* mel_val,
    int32_t n_mels, int32_t normalize, float* out, void* stream) {
    return emph::launch_logmel<float>(
        audio, audio_off, audio_len, chunk_start, chunk_len, row_start, n_seq,
        row_seq, total_rows, mel_ptr, mel_col, mel_val, n_mels, normalize, out, stream);
}

int emph_logmel_i16(
    const int16_t* audio,
    const int64_t* audio_off, const int32_t* audio_len,
    const int32_t* chunk_start, const int32_t* chunk_len,
    const int32_t* row_start, int32_t n_seq,
    const int32_t* row_seq, int32_t total_rows,
    const int32_t* mel_ptr, const int16_t* mel_col, const float* mel_val,
    int32_t n_mels, int32_t normalize, float* out, void* stream) {
    return emph::launch_logmel<int16_t>(
        audio, audio_off, audio_len, chunk_start, chunk_len, row_start, n_seq,
        row_seq, total_rows, mel_ptr, mel_col, mel_val, n_mels, normalize, out, stream);
}

}  // extern "C"
