// Fused log-mel feature extraction for sm_100a.
//
// Replaces, in ONE kernel and with each audio sample read from HBM once
// (overlapping frames are served from a shared-memory stage), the reference
// call chain
//   F.pad(zeros 432)            emphases/core.py:357-358
//   audio[:, s:e] chunk slice   emphases/core.py:395-401
//   F.pad(reflect 432)          emphases/data/preprocess/mels.py:32-36
//   torch.stft(1024, hop 160, hann, center=False)      mels.py:39-48
//   sqrt(re^2 + im^2 + 1e-6)    mels.py:51
//   basis @ spectrogram, log(clamp(., 1e-5)), optional (x+10)/10
//                               mels.py:94-109, 57-58
//
// Persistent CTAs of 8 warps, two per SM (so one CTA's mel / store phases
// overlap the other's FFT phase), each working on tiles of 16 consecutive
// packed rows:
//   phase 1  one warp per frame: 1024-point real FFT as a 512-point complex
//            FFT of the even/odd packed signal -- 3 radix-8 Stockham passes
//            on PACKED fp32 (FADD2 / FMUL2 / FFMA2: one warp instruction per
//            complex add or scale), 2 butterflies per lane per pass, twiddles
//            and the Hann window held in registers, two exchanges through
//            bank-padded shared memory, then the real-FFT unpacking done in
//            registers and the magnitudes written to a [16 frames][513 bins]
//            tile;
//   phase 2  mel projection with lane % 16 = frame.  For the band structure of
//            the default filterbank (every bin feeds the falling slope of one
//            row and the rising slope of the next) the projection is
//            straight-line generated code (mel_sweep.inc): each warp walks its
//            share of the spectrum once, one half-warp summing the falling
//            slopes and the other the rising slopes of the same 16 frames, so
//            every magnitude is read once (a 16-byte load both halves share)
//            and costs one FFMA per lane.  Any other basis takes a banded
//            dense table (or, if that does not fit, the CSR arrays in global
//            memory);
//   phase 3  log / clamp and a fully coalesced store of the 16 x n_mels tile.
// All fp32.  Bound: the shared-memory pipe and the FP32 pipe (about 25 kFLOP
// per frame against 960 B of HBM traffic), see DESIGN.md.
#include <cstddef>

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
constexpr int kRound = 128;               // tile descriptors are computed 128 tiles ahead
constexpr int kPartStride = 17;           // floats per mel row of the partial-sum tile

// The register-resident window (see the kernel) leaves the spectrum scaled by
// 4: the epsilon under the square root and the mel weights absorb it (powers
// of two: exact).
constexpr float kMagEps = 16.f * 1e-6f;
constexpr float kMelScale = 0.25f;

#include "mel_sweep.inc"
static_assert(kSweepWarps == kWarps, "regenerate mel_sweep.inc");

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

// What the default-structure mel path keeps in shared memory.  It overlays the
// general path's banded table (a launch uses one path or the other).
struct SweepSmem {
    float w[2][kMagStride];                        // falling / rising weight per bin
    // The two halves of a warp store part[j][f] and rising[j + 1][f] with one
    // instruction: the padding puts the two targets 16 banks apart (checked below).
    float bank_pad[19];
    float rising[kSweepMels + 2][kPartStride];     // rising-slope sums, row m + 1
};

struct __align__(16) LogmelSmem {
    float mag[kTile][kMagStride];  // bins 513..515 stay zero (band padding reads them)
    float2 xchg[kWarps][kXchg];
    // phase 2 -> phase 3, row m + 1: the falling-slope sums of the default path
    // (the rising ones are in sweep.rising), the whole sums of the general path
    float part[kMaxMels + 2][kPartStride];
    union {
        SweepSmem sweep;
        float4 mel_quad[kMaxQuads];    // general path: banded dense weights, 4 bins per entry
    };
    int32_t mel_first[kMaxMels];   // first quad of the row in mel_quad
    int32_t mel_bin0[kMaxMels];    // first bin of the row's band (multiple of 4)
    int32_t mel_quads[kMaxMels / 2];   // double-quads per row of the pair (both rows padded to it)
    int32_t mel_fits;              // 0: the basis does not fit the table, use the CSR path
    int32_t sweep_ok;              // 1: the basis has the structure mel_sweep.inc was generated for
    // audio of one interior tile (16 frames = 3424 samples), fetched with one
    // cp.async.bulk while the previous tile is in its mel / store phases
    __align__(16) unsigned char stage[kSpan * sizeof(float)];
    long long tile_src[kRound];    // per tile of this CTA: element index of the span, or -1
    unsigned long long bar;        // mbarrier the bulk copy completes on
};
static_assert(sizeof(SweepSmem) <= sizeof(float4) * kMaxQuads, "sweep tables must fit the overlay");
static_assert(2 * (sizeof(LogmelSmem) + 1024) <= 233472, "two CTAs per SM");
static_assert(
    ((offsetof(LogmelSmem, sweep) + offsetof(SweepSmem, rising) - offsetof(LogmelSmem, part)) / 4 +
     kPartStride) % 32 == 16, "falling / rising stores must hit disjoint banks");

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


// ---- packed fp32 arithmetic (sm_100a FADD2 / FMUL2 / FFMA2) ----
// A complex number lives in an aligned register pair (x = re, y = im) and is
// moved through the FMA pipe by ONE warp instruction per complex add / scale.
// ptxas folds a component swap, a per-component negation and a scalar
// broadcast into the operand modifiers of the packed instruction
// (R.F32x2.LO_HI.NP, -R.F32x2.HI_LO, R.F32), so multiplying by -i, +i, conj
// and a real factor costs no instruction of its own: the helpers below build
// those operands as plain C2 values.
struct C2 { float x, y; };

__device__ __forceinline__ C2 pk_add(C2 a, C2 b) {
    C2 r;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ C2 pk_sub(C2 a, C2 b) {
    C2 r;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "sub.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ C2 pk_mul(C2 a, C2 b) {
    C2 r;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ C2 pk_fma(C2 a, C2 b, C2 c) {
    C2 r;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "mov.b64 rc, {%6, %7};\n\tfma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
__device__ __forceinline__ C2 mul_mi(C2 a) { return C2{a.y, -a.x}; }    // -i a
__device__ __forceinline__ C2 mul_pi(C2 a) { return C2{-a.y, a.x}; }    // +i a
__device__ __forceinline__ C2 conj2(C2 a) { return C2{a.x, -a.y}; }
__device__ __forceinline__ C2 bcast(float s) { return C2{s, s}; }
// a * (w.x + i w.y): two packed instructions
__device__ __forceinline__ C2 cmul(C2 a, float2 w) {
    return pk_fma(a, bcast(w.x), pk_mul(mul_pi(a), bcast(w.y)));
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

// 8-point DFT after its first butterfly stage: a[r] = v[r] + v[r + 4],
// a[r + 4] = v[r] - v[r + 4] (r < 4) in, natural-order X[0..7] out.
// 18 packed instructions; the two 1/sqrt(2) rotations ride on the FFMA2s of
// the last stage.
__device__ __forceinline__ void fft8_tail(C2 (&a)[8]) {
    const float s = 0.70710678118654752440f;
    // even half: 4-point DFT of a[0..3] -> X0 X2 X4 X6
    const C2 b0 = pk_add(a[0], a[2]), b2 = pk_sub(a[0], a[2]);
    const C2 b1 = pk_add(a[1], a[3]), b3 = pk_sub(a[1], a[3]);
    // odd half: 4-point DFT of (a4, a5 w8, a6 w8^2, a7 w8^3) -> X1 X3 X5 X7
    const C2 c4 = pk_add(a[4], mul_mi(a[6])), c6 = pk_add(a[4], mul_pi(a[6]));
    const C2 t5 = pk_add(a[5], mul_mi(a[5]));          // a5 (1 - i)   = sqrt2 a5 w8
    const C2 t7 = pk_sub(mul_mi(a[7]), a[7]);          // -a7 (1 + i)  = sqrt2 a7 w8^3
    const C2 c5 = pk_add(t5, t7), c7 = pk_sub(t5, t7);
    a[0] = pk_add(b0, b1);
    a[4] = pk_sub(b0, b1);
    a[2] = pk_add(b2, mul_mi(b3));
    a[6] = pk_add(b2, mul_pi(b3));
    a[1] = pk_fma(c5, bcast(s), c4);
    a[5] = pk_fma(c5, bcast(-s), c4);
    a[3] = pk_fma(mul_mi(c7), bcast(s), c6);
    a[7] = pk_fma(mul_mi(c7), bcast(-s), c6);
}
// 8-point DFT in place; v[] ends up in natural order (26 packed instructions)
__device__ __forceinline__ void fft8(C2 (&v)[8]) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const C2 t = v[r];
        v[r] = pk_add(t, v[r + 4]);
        v[r + 4] = pk_sub(t, v[r + 4]);
    }
    fft8_tail(v);
}

template <typename T>
__device__ __forceinline__ float to_float(T v);
template <>
__device__ __forceinline__ float to_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_float<int16_t>(int16_t v) {
    return (float)v * (1.f / 32768.f);
}

// Sample-rate conversion fused into the front end (emphases/core.py:613-619:
// the reference resamples with torchaudio.transforms.Resample before anything
// else).  With it the packed `audio` is at the SOURCE rate and sample n of an
// utterance at the model rate is the polyphase windowed-sinc sum
//   y[q * fresh + p] = sum_k filter[p][k] * x[q * orig - width + k]
// over the taps whose filter value is not zero (tap_lo[p] .. tap_hi[p]), the
// same sum in the same order as csrc/resample.cu -- bit-identical samples,
// which then never exist in HBM.
struct Resampler {
    const float* filter;          // [fresh][taps], resampling.filter_bank
    const int32_t* tap_lo;        // [fresh] first tap with a non-zero value
    const int32_t* tap_hi;        // [fresh] one past the last one
    const int32_t* source_len;    // [n_seq] samples of the sequence's utterance at the source rate
    int orig, fresh, width, taps;
};

template <typename T>
__device__ __forceinline__ float resampled_sample(
    const Resampler& rs, const T* __restrict__ src, int src_len, int n) {
    const int q = n / rs.fresh, p = n - q * rs.fresh;
    const float* __restrict__ w = rs.filter + (size_t)p * rs.taps;
    const long long first = (long long)q * rs.orig - rs.width;      // source index of tap 0
    long long k0 = __ldg(rs.tap_lo + p), k1 = __ldg(rs.tap_hi + p);
    if (-first > k0) k0 = -first;
    if ((long long)src_len - first < k1) k1 = (long long)src_len - first;
    float acc = 0.f;
    for (long long k = k0; k < k1; ++k)
        acc = fmaf(__ldg(w + k), to_float<T>(__ldg(src + first + k)), acc);
    return acc;
}

// Sample q of the reflect-padded chunk, through the whole index map
// (SURVEY.md A.2): chunk[j] = P[s + j], P = zeros(432) ++ audio ++ zeros(432)
template <typename T, bool RESAMPLE>
__device__ __forceinline__ float chunk_sample(
    const T* __restrict__ audio, int T_len, int s, int L, int q, const Resampler& rs,
    int src_len) {
    int j = q - kPad;
    if (j < 0) j = -j;
    if (j >= L) j = 2 * (L - 1) - j;
    int a = s + j - kPad;
    if (!(a >= 0 && a < T_len)) return 0.f;
    if constexpr (RESAMPLE) return resampled_sample<T>(rs, audio, src_len, a);
    else return to_float<T>(audio[a]);
}

template <typename T, bool RESAMPLE>
__global__ void __launch_bounds__(kWarps * 32, 2)
logmel_kernel(
    const T* __restrict__ audio,
    const int64_t* __restrict__ audio_off, const int32_t* __restrict__ audio_len,
    const int32_t* __restrict__ chunk_start, const int32_t* __restrict__ chunk_len,
    const int32_t* __restrict__ row_start,
    const int32_t* __restrict__ row_seq, int total_rows,
    const int32_t* __restrict__ mel_ptr, const int16_t* __restrict__ mel_col,
    const float* __restrict__ mel_val, int n_mels, int normalize,
    float* __restrict__ out, Resampler rs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LogmelSmem& sm = *reinterpret_cast<LogmelSmem*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- mel tables, built once per CTA from the CSR basis ----
    for (int i = tid; i < kTile * (kMagStride - kBins); i += blockDim.x)
        sm.mag[i / (kMagStride - kBins)][kBins + i % (kMagStride - kBins)] = 0.f;
    for (int i = tid; i < kMaxQuads; i += blockDim.x)      // also zeroes the sweep overlay
        sm.mel_quad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < (kMaxMels + 2) * kPartStride; i += blockDim.x)
        (&sm.part[0][0])[i] = 0.f;
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
    // Default structure (mel_sweep.inc): every non-zero of row m sits in a bin of
    // segment m (rising slope) or m + 1 (falling slope).  The weights themselves
    // are taken from the caller's basis.
    int misfit = n_mels != kSweepMels;
    if (!misfit) {
        for (int m = warp; m < n_mels; m += kWarps) {
            for (int e = mel_ptr[m] + lane; e < mel_ptr[m + 1]; e += 32) {
                const int k = mel_col[e];
                const float v = mel_val[e];
                if (v == 0.f) continue;
                const int seg = (k >= 0 && k < kBins) ? kSweepBinSeg[k] : -2;
                if (seg == m) sm.sweep.w[1][k] = v * kMelScale;
                else if (seg == m + 1) sm.sweep.w[0][k] = v * kMelScale;
                else misfit = 1;
            }
        }
    }
    misfit = __syncthreads_or(misfit);
    if (tid == 0) {
        sm.sweep_ok = misfit ? 0 : 1;
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
    const bool sweep = sm.sweep_ok != 0;
    const bool banded = sm.mel_fits != 0;
    if (!sweep) {       // the overlay holds half-written sweep weights: clear it
        for (int i = tid; i < kMaxQuads; i += blockDim.x)
            sm.mel_quad[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
    }
    if (!sweep && banded) {
        float* const table = reinterpret_cast<float*>(sm.mel_quad);
        for (int m = warp; m < n_mels; m += kWarps) {
            const int base = 4 * sm.mel_first[m] - sm.mel_bin0[m];
            for (int e = mel_ptr[m] + lane; e < mel_ptr[m + 1]; e += 32)
                table[base + mel_col[e]] = mel_val[e] * kMelScale;
        }
    }

    // Per-lane twiddles, kept in registers for the whole kernel:
    //   pass 2: exp(-2 pi i * 8 r (lane % 8) / 512)
    //   pass 3: exp(-2 pi i r j / 512) for butterfly j = lane.  The lane's second
    //           butterfly is j' = 64 - lane, whose twiddle is
    //           conj(tw3[r]) * exp(-2 pi i r / 8): the conjugate is an operand
    //           modifier and the second factor only rotates the outputs of the
    //           8-point DFT by one slot, so it needs no registers of its own.
    //           Lane 0 works on the self-paired residues j = 0 (unit twiddles,
    //           its multiplications are predicated off) and j' = 32 (same rule
    //           with tw3[r] = exp(-2 pi i 32 r / 512)).
    //   unpack: exp(-2 pi i k / 1024), k the bin of slot q
    float2 tw2[8], tw3[8], wq[8];
#pragma unroll
    for (int r = 1; r < 8; ++r) {
        float sn, cs;
        sincospif(-2.f * (float)(8 * r * (lane & 7)) / (float)kHalf, &sn, &cs);
        tw2[r] = make_float2(cs, sn);
        sincospif(-2.f * (float)(r * (lane ? lane : 32)) / (float)kHalf, &sn, &cs);
        tw3[r] = make_float2(cs, sn);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        // bin of unpacking slot q: lane + 64 q; lane 0 takes 64 q (q < 4) and 32 + 64 (q - 4)
        const int k = (lane == 0 && q >= 4) ? 64 * q - 224 : lane + 64 * q;
        sincospif(-2.f * (float)k / (float)kFft, &wq[q].y, &wq[q].x);
    }
    // Window in registers.  The first radix-8 stage pairs sample n with
    // n + 512, and the periodic Hann window satisfies
    //   w[n] = 1/2 - 1/2 cos(2 pi n / 1024),  w[n + 512] = 1/2 + 1/2 cos(..),
    // so with s = x[n] + x[n + 512], d = x[n] - x[n + 512] and C = cos(..):
    //   x[n] w[n] + x[n + 512] w[n + 512] = (s - C d) / 2
    //   x[n] w[n] - x[n + 512] w[n + 512] = (d - C s) / 2
    // i.e. window AND first butterfly in 2 FADD2 + 2 FFMA2 per complex pair
    // from 16 per-lane cosines (samples 4 lane + 128 r + c, r < 4) -- no window
    // table in shared memory.  The 1/2, the 1/2 of the real-FFT unpacking and
    // the 1/4 they leave on the magnitudes are folded into kMagEps and the mel
    // weights (all powers of two: exact).
    C2 cw0[4], cw1[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int n = 4 * lane + 128 * r;
        cw0[r] = C2{-cospif(2.f * (float)n / (float)kFft), -cospif(2.f * (float)(n + 1) / (float)kFft)};
        cw1[r] = C2{-cospif(2.f * (float)(n + 2) / (float)kFft), -cospif(2.f * (float)(n + 3) / (float)kFft)};
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
                    (RESAMPLE || (reinterpret_cast<uintptr_t>(audio + index) & 15) == 0);
                // (resampling variant: the tile's samples are computed, not copied:
                // sequence and first model-rate sample)
                if (ok) desc = RESAMPLE ? ((long long)u << 32) | (unsigned)a0 : index;
            }
        }
        if (tid < kRound) sm.tile_src[tid] = desc;
    };
    if (tid == 0) mbar_init(&sm.bar, 1);
    describe_round(0);
    __syncthreads();
    if (!RESAMPLE && tid == 0 && sm.tile_src[0] >= 0)
        bulk_fetch(sm.stage, audio + sm.tile_src[0], kSpan * sizeof(T), &sm.bar);
    uint32_t parity = 0;
    int it = 0;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int row0 = tile * kTile;
        const bool staged = sm.tile_src[it & (kRound - 1)] >= 0;
        if (staged) {
            if constexpr (RESAMPLE) {
                // the tile's 3424 model-rate samples, resampled cooperatively into
                // the stage (the previous tile's FFT phase is long over)
                const long long desc = sm.tile_src[it & (kRound - 1)];
                const int u = (int)(desc >> 32), a0 = (int)(desc & 0xffffffffLL);
                const T* src = audio + __ldg(audio_off + u);
                const int src_len = __ldg(rs.source_len + u);
                float* stage = reinterpret_cast<float*>(sm.stage);
                for (int i = tid; i < kSpan; i += blockDim.x)
                    stage[i] = resampled_sample<T>(rs, src, src_len, a0 + i);
                __syncthreads();
            } else {
                mbar_wait(&sm.bar, parity);
                parity ^= 1;
            }
        }

        // ======================= phase 1: FFT + magnitude =======================
#pragma unroll 1
        for (int f = warp; f < kTile; f += kWarps) {
            // ---- load 1024 samples as 512 complex ----
            // pass 1: lane handles butterflies j = 2 lane (v0) and 2 lane + 1 (v1);
            // inputs z[j + 64 r], so one 16-byte load brings both butterflies' inputs
            C2 v0[8], v1[8];
            auto load_run = [&](auto* p) {    // 1024 contiguous samples
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int n = 4 * lane + 128 * r;
                    if constexpr (sizeof(*p) == 4) {
                        const float4 x = *reinterpret_cast<const float4*>(p + n);
                        v0[r] = C2{x.x, x.y};
                        v1[r] = C2{x.z, x.w};
                    } else {
                        const short4 x = *reinterpret_cast<const short4*>(p + n);
                        v0[r] = C2{to_float<int16_t>(x.x), to_float<int16_t>(x.y)};
                        v1[r] = C2{to_float<int16_t>(x.z), to_float<int16_t>(x.w)};
                    }
                }
            };
            if (staged) {
                if constexpr (RESAMPLE)
                    load_run(reinterpret_cast<const float*>(sm.stage) + f * kHop);
                else
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
                const bool interior = !RESAMPLE && (q0 >= kPad) && (q0 + kFft - kPad <= L) &&
                                      (a0 >= 0) && (a0 + kFft <= T_len);
                const int src_len = RESAMPLE ? __ldg(rs.source_len + u) : 0;
                if (interior) {
                    load_run(src + a0);
                } else {
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const int n = q0 + 4 * lane + 128 * r;
                        v0[r] = C2{
                            chunk_sample<T, RESAMPLE>(src, T_len, s, L, n, rs, src_len),
                            chunk_sample<T, RESAMPLE>(src, T_len, s, L, n + 1, rs, src_len)};
                        v1[r] = C2{
                            chunk_sample<T, RESAMPLE>(src, T_len, s, L, n + 2, rs, src_len),
                            chunk_sample<T, RESAMPLE>(src, T_len, s, L, n + 3, rs, src_len)};
                    }
                }
            }
            // pass 1: window + first butterfly stage (see cw0 / cw1), Ns = 1, no
            // twiddles, out[8 j + r]
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const C2 s0 = pk_add(v0[r], v0[r + 4]), d0 = pk_sub(v0[r], v0[r + 4]);
                const C2 s1 = pk_add(v1[r], v1[r + 4]), d1 = pk_sub(v1[r], v1[r + 4]);
                v0[r] = pk_fma(cw0[r], d0, s0);
                v0[r + 4] = pk_fma(cw0[r], s0, d0);
                v1[r] = pk_fma(cw1[r], d1, s1);
                v1[r + 4] = pk_fma(cw1[r], s1, d1);
            }
            fft8_tail(v0);
            fft8_tail(v1);
#pragma unroll
            for (int r = 0; r < 8; ++r)     // slots 66 r + 2 lane, + 1: one 16-byte store
                *reinterpret_cast<float4*>(pw + lane + 66 * r) =
                    make_float4(v0[r].x, v0[r].y, v1[r].x, v1[r].y);
            __syncwarp();

            // pass 2: Ns = 8; twiddle exp(-2 pi i r (j % 8) / 64)
            {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float2 x0 = pr2[8 * r], x1 = pr2[4 + 8 * r];
                    v0[r] = C2{x0.x, x0.y};
                    v1[r] = C2{x1.x, x1.y};
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
                    pw[40 * (r & 3) + 304 * (r >> 2)] = make_float2(v0[r].x, v0[r].y);
                    pw[152 + 40 * (r & 3) + 304 * (r >> 2)] = make_float2(v1[r].x, v1[r].y);
                }
                __syncwarp();
            }

            // pass 3: Ns = 64; twiddle exp(-2 pi i r j / 512), results stay in
            // registers: v0[r] = Z[lane + 64 r], v1[(r + 1) % 8] = Z[j1 + 64 r]
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float2 x0 = pr3[8 * (r & 3) + 152 * (r >> 2)];
                const float2 x1 = pr3b[8 * (r & 3) + 152 * (r >> 2)];
                v0[r] = C2{x0.x, x0.y};
                v1[r] = C2{x1.x, x1.y};
            }
            __syncwarp();                      // xw is free for the next frame
            if (lane != 0) {
#pragma unroll
                for (int r = 1; r < 8; ++r) v0[r] = cmul(v0[r], tw3[r]);
            }
#pragma unroll
            for (int r = 1; r < 8; ++r) v1[r] = cmul(v1[r], make_float2(tw3[r].x, -tw3[r].y));
            fft8(v0);
            fft8(v1);

            // ---- real-FFT unpacking in registers ----
            // With e = (Z[k] + conj Z[512-k]) / 2, d = (Z[k] - conj Z[512-k]) / 2,
            // o = -i d (the 1/2 is already in the window) and
            // w = exp(-2 pi i k / 1024):
            // X[k] = e + w o,  X[512-k] = conj(e - w o),
            // so one (e, w o) serves both bins of a pair: 6 packed instructions.
            // (The magnitudes are taken from the complex sums, not as
            // |e|^2 + |o|^2 +- 2 Re(e conj(w o)): bins k and 512-k of speech
            // differ by up to 60 dB and subtracting powers would wipe out the
            // weak one.)
            // With zb(r) = v1[(r + 1) % 8] = Z[j1 + 64 r]: slot q of a lane != 0
            // pairs k = lane + 64 q (v0[q]) with 512 - k = (64 - lane) + 64 (7 - q)
            // (zb(7 - q)): no shuffles.  Lane 0 holds the self-paired residues 0
            // and 32: slots 0..3 pair v0[q] with v0[(8 - q) % 8] (k = 64 q), slots
            // 4..7 pair zb(q - 4) with zb(11 - q) (k = 32 + 64 (q - 4)); v0[4] =
            // Z[256] is its own partner (below).
            float* const mag = sm.mag[f];
            const bool lane0 = lane == 0;
            float* const mag_lo = mag + lane;                       // slots 0..3: mag_lo[64 q]
            float* const mag_hi = mag + (lane0 ? -224 : lane);      // slots 4..7: mag_hi[64 q]
            float* const mir_lo = mag - lane;                       // mir_lo[512 - 64 q]
            float* const mir_hi = mag - (lane0 ? -224 : lane);      // mir_hi[512 - 64 q]
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                C2 a = v0[q], pz = v1[(8 - q) & 7];                  // zb(7 - q)
                if (q < 4) {
                    const C2 alt = v0[(8 - q) & 7];
                    pz.x = lane0 ? alt.x : pz.x;
                    pz.y = lane0 ? alt.y : pz.y;
                } else {
                    const C2 alt_a = v1[(q - 3) & 7], alt = v1[(12 - q) & 7];   // zb(q - 4), zb(11 - q)
                    a.x = lane0 ? alt_a.x : a.x;
                    a.y = lane0 ? alt_a.y : a.y;
                    pz.x = lane0 ? alt.x : pz.x;
                    pz.y = lane0 ? alt.y : pz.y;
                }
                const C2 e = pk_add(a, conj2(pz));
                const C2 d = pk_sub(a, conj2(pz));
                // w o = w (-i d) = (-i d) w.x + d w.y   [i (-i d) = d]
                const C2 wo = pk_fma(mul_mi(d), bcast(wq[q].x), pk_mul(d, bcast(wq[q].y)));
                const C2 x = pk_add(e, wo), y = pk_sub(e, wo);
                (q < 4 ? mag_lo : mag_hi)[64 * q] = sqrt_pos(fmaf(x.x, x.x, fmaf(x.y, x.y, kMagEps)));
                (q < 4 ? mir_lo : mir_hi)[kHalf - 64 * q] = sqrt_pos(fmaf(y.x, y.x, fmaf(y.y, y.y, kMagEps)));
            }
            // k = 256 pairs with itself: X[256] = conj(Z[256]); Z[256] = Z_8 of lane 0
            // (Z is X / 2: |X|^2 = 4 |Z|^2)
            if (lane0) mag[kHalf / 2] = sqrt_pos(fmaf(4.f * v0[4].x, v0[4].x, fmaf(4.f * v0[4].y, v0[4].y, kMagEps)));
        }
        __syncthreads();

        // every warp is done with the staged audio: fetch the next tile's span
        // (it lands while this tile is in its mel / store phases)
        if (((it + 1) & (kRound - 1)) == 0) {
            describe_round(it + 1);
            __syncthreads();
        }
        if (!RESAMPLE && tid == 0 && tile + (int)gridDim.x < n_tiles) {
            const long long next = sm.tile_src[(it + 1) & (kRound - 1)];
            if (next >= 0) bulk_fetch(sm.stage, audio + next, kSpan * sizeof(T), &sm.bar);
        }

        // ============== phase 2: mel projection, lane % 16 = frame ==============
        if (sweep) {
            // half-warp h sums slope h of every segment: segment j is row j - 1 of
            // the falling table and row j of the rising one (rows are stored + 1)
            const int h = lane >> 4, f = lane & 15;
            mel_sweep_default<kPartStride>(
                warp, reinterpret_cast<const float4*>(sm.mag[f]),
                reinterpret_cast<const float4*>(sm.sweep.w[h]),
                h ? &sm.sweep.rising[1][f] : &sm.part[0][f]);
        } else {
            // general basis: a half-warp per mel row, 16 frames at a time
            const int h = lane >> 4;                   // which row of the pair
            const int n_pairs = (n_mels + 1) >> 1;
            for (int half = 0; half < kTile / 16; ++half) {
                const int f = 16 * half + (lane & 15);
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
                        acc *= kMelScale;
                    }
                    sm.part[m + 1][f] = acc;
                }
            }
        }
        __syncthreads();

        // =============== phase 3: log, clamp and coalesced store ===============
#pragma unroll 1
        for (int f = warp; f < kTile; f += kWarps) {
            const int row = row0 + f;
            if (row >= total_rows) break;
            const bool live = staged || __ldg(row_seq + row) >= 0;
            float* const dst = out + (size_t)row * n_mels;
            for (int m = lane; m < n_mels; m += 32) {
                float v = sm.part[m + 1][f];
                if (sweep) v += sm.sweep.rising[m + 1][f];
                v = log_pos(fmaxf(v, 1e-5f));
                if (normalize) v = (v + 10.f) / 10.f;
                dst[m] = live ? v : 0.f;
            }
        }
        // the next tile's phase 1 only touches mag / xchg; part is rewritten
        // after the next __syncthreads
    }
}

template <typename T, bool RESAMPLE>
int launch_logmel(
    const T* audio,
    const int64_t* audio_off, const int32_t* audio_len,
    const int32_t* chunk_start, const int32_t* chunk_len,
    const int32_t* row_start, int32_t n_seq,
    const int32_t* row_seq, int32_t total_rows,
    const int32_t* mel_ptr, const int16_t* mel_col, const float* mel_val,
    int32_t n_mels, int32_t normalize, float* out, void* stream,
    Resampler rs = Resampler{nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0}) {
    EMPH_REQUIRE(n_seq >= 0 && total_rows >= 0, "emph_logmel: negative size");
    EMPH_REQUIRE(n_mels > 0 && n_mels <= kMaxMels, "emph_logmel: n_mels %d out of range", n_mels);
    if (total_rows == 0) return EMPH_OK;
    const size_t smem = sizeof(LogmelSmem);
    int s = check_cuda(
        cudaFuncSetAttribute(
            logmel_kernel<T, RESAMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
        "logmel smem attribute");
    if (s != EMPH_OK) return s;
    const int n_tiles = (total_rows + kTile - 1) / kTile;
    const int grid = n_tiles < 2 * sm_count() ? n_tiles : 2 * sm_count();
    logmel_kernel<T, RESAMPLE><<<grid, kWarps * 32, smem, (cudaStream_t)stream>>>(
        audio, audio_off, audio_len, chunk_start, chunk_len, row_start,
        row_seq, total_rows, mel_ptr, mel_col, mel_val, n_mels, normalize, out, rs);
    EMPH_CHECK_LAUNCH("emph_logmel");
    return EMPH_OK;
}

template <typename T>
int launch_logmel_resampled(
    const T* audio, const int64_t* audio_off, const int32_t* source_len,
    const int32_t* audio_len, const int32_t* chunk_start, const int32_t* chunk_len,
    const int32_t* row_start, int32_t n_seq, const int32_t* row_seq, int32_t total_rows,
    const int32_t* mel_ptr, const int16_t* mel_col, const float* mel_val,
    int32_t n_mels, int32_t normalize,
    const float* filter, const int32_t* tap_lo, const int32_t* tap_hi,
    int32_t orig_freq, int32_t new_freq, int32_t width, float* out, void* stream) {
    EMPH_REQUIRE(filter && tap_lo && tap_hi && source_len, "emph_logmel_resampled: null filter");
    EMPH_REQUIRE(orig_freq > 0 && new_freq > 0 && width >= 0, "emph_logmel_resampled: bad filter");
    const Resampler rs{filter, tap_lo, tap_hi, source_len, orig_freq, new_freq, width,
                       2 * width + orig_freq};
    return launch_logmel<T, true>(
        audio, audio_off, audio_len, chunk_start, chunk_len, row_start, n_seq, row_seq,
        total_rows, mel_ptr, mel_col, mel_val, n_mels, normalize, out, stream, rs);
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
    return emph::launch_logmel<float, false>(
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
    return emph::launch_logmel<int16_t, false>(
        audio, audio_off, audio_len, chunk_start, chunk_len, row_start, n_seq,
        row_seq, total_rows, mel_ptr, mel_col, mel_val, n_mels, normalize, out, stream);
}

int emph_logmel_resampled_f32(
    const float* audio, const int64_t* audio_off, const int32_t* source_len,
    const int32_t* audio_len, const int32_t* chunk_start, const int32_t* chunk_len,
    const int32_t* row_start, int32_t n_seq, const int32_t* row_seq, int32_t total_rows,
    const int32_t* mel_ptr, const int16_t* mel_col, const float* mel_val,
    int32_t n_mels, int32_t normalize,
    const float* filter, const int32_t* tap_lo, const int32_t* tap_hi,
    int32_t orig_freq, int32_t new_freq, int32_t width, float* out, void* stream) {
    return emph::launch_logmel_resampled<float>(
        audio, audio_off, source_len, audio_len, chunk_start, chunk_len, row_start, n_seq,
        row_seq, total_rows, mel_ptr, mel_col, mel_val, n_mels, normalize, filter, tap_lo,
        tap_hi, orig_freq, new_freq, width, out, stream);
}

int emph_logmel_resampled_i16(
    const int16_t* audio, const int64_t* audio_off, const int32_t* source_len,
    const int32_t* audio_len, const int32_t* chunk_start, const int32_t* chunk_len,
    const int32_t* row_start, int32_t n_seq, const int32_t* row_seq, int32_t total_rows,
    const int32_t* mel_ptr, const int16_t* mel_col, const float* mel_val,
    int32_t n_mels, int32_t normalize,
    const float* filter, const int32_t* tap_lo, const int32_t* tap_hi,
    int32_t orig_freq, int32_t new_freq, int32_t width, float* out, void* stream) {
    return emph::launch_logmel_resampled<int16_t>(
        audio, audio_off, source_len, audio_len, chunk_start, chunk_len, row_start, n_seq,
        row_seq, total_rows, mel_ptr, mel_col, mel_val, n_mels, normalize, filter, tap_lo,
        tap_hi, orig_freq, new_freq, width, out, stream);
}

}  // extern "C"
