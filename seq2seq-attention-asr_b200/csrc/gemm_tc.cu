// gemm_tc.cu -- tcgen05 / TMEM GEMM for the large time-batched projections:
//     C[M,N] = alpha * op(A) . op(B) (+ beta C) (+ bias[n])        fp32 in, fp32 out
// (the x-columns of the GRU gates over all frames, LinearZeroBias.lua:42; Vh = h W_V^T,
// TemporalConvolutionZeroBias.lua:39; their data and weight gradients.)
//
// Precision: the reference computes these products in fp32 BLAS and the parity bound is 1e-4, which a
// single TF32 pass (10-bit mantissa, ~5e-4) misses.  Each operand x is therefore split into
// hi = top 19 bits of x (exactly representable in TF32) and lo = x - hi, and three tensor-core products
// A_lo B_hi + A_hi B_lo + A_hi B_hi are accumulated in the fp32 TMEM accumulator ("3xTF32", error ~2^-21).
//
// The split is done ONCE per operand by a streaming pre-pass (split_kernel / transpose_split_kernel) that
// also brings M- or N-contiguous operands to the K-contiguous form the tensor core wants, so the GEMM kernel
// itself is a pure TMA -> tcgen05.mma pipeline (the first version split inside shared memory, which cost more
// shared-memory bandwidth than the MMAs themselves and re-split every slab once per tile that used it).
//
// Kernel: persistent, one CTA per SM, 128 x 256 output tiles, K consumed in 32-float (= 128-byte swizzle row)
// slabs, two 96 KB stages {A_hi, A_lo, B_hi, B_lo}:
//   warp 0      TMA producer: 4 cp.async.bulk.tensor.2d (SWIZZLE_128B) per slab
//   warp 1      TMEM allocation (512 columns = two 128x256 fp32 accumulators) and the single-thread
//               tcgen05.mma.cta_group::1.kind::tf32 issue loop (M=128, N=256, K=8; 12 instructions per slab);
//               tcgen05.commit frees the stage / publishes the accumulator
//   warps 2-9   epilogue: tcgen05.ld (32 lanes x 32 columns) -> smem transpose -> coalesced stores, overlapped
//               with the next segment's main loop through the second accumulator
// Scheduling: tiles are dealt round-robin to the CTAs; the tiles left over after the last full round are cut
// along K into equal shares ("stream-K" tail) whose partial sums are added with atomics, so that e.g. 150 tiles
// on 148 SMs cost ~1.25 tile times instead of 2.  Weight gradients (few tiles, K = B*L) are all tail.
#include <cuda.h>
#include <stdio.h>

#include "common.cuh"

namespace s2s {

namespace tc {
constexpr int BM = 128, BK = 32;                    // BK floats = 128 bytes = one swizzle row
constexpr int A_BYTES = BM * BK * 4;                // 16 KB
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int SCR_FLOATS = 32 * 33;                 // per-warp transpose scratch
// tile width BN in {64, 128, 256}: narrow outputs (64- / 128-plane convolutions, K = 123 weight gradients) do not pay for
// 256-wide MMAs, and their smaller stages buy a deeper pipeline
constexpr int b_bytes(int BN) { return BN * BK * 4; }
constexpr int stage_bytes(int BN) { return 2 * A_BYTES + 2 * b_bytes(BN); }        // A_hi, A_lo, B_hi, B_lo
constexpr int stages(int BN) { return BN == 256 ? 2 : (BN == 128 ? 3 : 4); }       // 192 KB of stages in every case
constexpr size_t smem_bytes(int BN) { return (size_t)stages(BN) * stage_bytes(BN) + (size_t)EPI_WARPS * SCR_FLOATS * 4 + 1024 /*align*/ + 256 /*barriers*/; }
constexpr int MIN_SHARE = 4;                        // a stream-K share is at least this many slabs

struct Sched {
    int tiles_m, tiles_n, nk;     // tile grid and slabs per tile
    int G;                        // CTAs
    int R;                        // full rounds: tiles [0, R*G) are done whole, tile i*G + g by CTA g
    int rem;                      // tiles left for the stream-K tail
    int Gp;                       // CTAs that take a share of the tail
    int dbg;                      // timing experiments (S2S_TC_DBG): 1 = no MMAs (TMA stream only), 2 = no TMA (MMA issue only)
    // implicit 3x3 convolution (channels-last activations on their full grid): K = taps * tap_slabs * 32; slab kb belongs to
    // tap kb / tap_slabs and reads the A operand tap_row[tap] ROWS further down (forward: +(kh*W + kw) pixels, data gradient:
    // minus that; rows outside the matrix read as zeros) at K offset (kb % tap_slabs) * 32.  b_col0: K offset added to the B
    // operand's coordinate (weight gradient of one tap: the transposed activations shifted by the tap's pixel offset).
    int tap_slabs;                // 0 = plain GEMM
    int tap_row[9];
    int b_col0;
    int relu;                     // C = max(., 0) in the epilogue (whole tiles only: no atomic tail)
    // MN-major operands (the NN data-gradient and TN weight-gradient forms): the operand lies in memory as [K, MN] (MN contiguous) and is
    // fetched untransposed, one 2-D TMA box {32 MN-floats, 32 K-rows} per 32-wide MN chunk; shared memory then holds, per chunk, 32 K-rows
    // of 128 bytes (4 KB, 128-byte swizzle with 32-byte atoms: the only MN-major layout the tensor core takes for 32-bit operands, see
    // make_desc_mn).  Columns / rows beyond the matrix read as zeros.
    int a_mn, b_mn;
};
struct Span { int tm, tn, kb0, kb1; };

// i-th segment of CTA g; false when the CTA has no more work.  All three roles enumerate the same sequence.
__device__ __forceinline__ bool get_seg(const Sched& s, int g, int i, Span& seg) {
    int tile, kb0, kb1;
    if (i < s.R) {
        tile = i * s.G + g; kb0 = 0; kb1 = s.nk;
    } else {
        if (s.rem == 0 || g >= s.Gp) return false;
        const long long U = (long long)s.rem * s.nk;
        const long long u0 = U * g / s.Gp, u1 = U * (g + 1) / s.Gp;
        if (u1 <= u0) return false;
        const int t0 = (int)(u0 / s.nk);
        const int t = t0 + (i - s.R);
        if ((long long)t * s.nk >= u1) return false;
        const long long lo = (long long)t * s.nk;
        kb0 = (int)((u0 > lo ? u0 : lo) - lo);
        const long long hi = lo + s.nk;
        kb1 = (int)((u1 < hi ? u1 : hi) - lo);
        tile = s.R * s.G + t;
    }
    seg.tm = tile / s.tiles_n; seg.tn = tile - seg.tm * s.tiles_n; seg.kb0 = kb0; seg.kb1 = kb1;
    return true;
}

__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, unsigned parity) {
    // a mis-programmed pipeline must fail loudly instead of hanging the GPU
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
// MN-major operand tile of `chunks` x 32 MN-columns starting at column mn0, K rows [k0, k0 + 32)
__device__ __forceinline__ void tma_load_mn(unsigned char* smem_dst, const CUtensorMap* map, int mn0, int k0, int chunks, uint64_t* bar) {
    for (int c = 0; c < chunks; c++) tma_load_2d(smem_dst + c * 4096, map, mn0 + 32 * c, k0, bar);
}
// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);               // start address
    d |= (uint64_t)0 << 16;                                // leading-dimension byte offset (unused: swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                      // stride-dimension byte offset
    d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
    return d;
}
// MN-major operand (see Sched::a_mn).  For 32-bit (tf32) MN-major operands the tensor core accepts ONE swizzled layout: "128-byte swizzle
// with 32-byte atomicity" (layout type 1): rows of 128 bytes = 32 MN-floats of one K index, the four 32-byte pieces of a row XOR-ed with
// (K index & 3) -- what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  The swizzle atom is 32 MN x 4 K (512 bytes):
// stride offset = 512 (one group of 4 K-rows to the next), leading offset = 4096 (one 32-wide MN chunk to the next).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(4096 >> 4) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                                // SWIZZLE_128B_BASE32B
    return d;
}
// instruction descriptor: D = f32, A = B = tf32, N = BN, M = 128; bits 15 / 16: A / B operand is MN-major (0 = K-major)
template <int BN>
__device__ __forceinline__ uint32_t make_idesc() {
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format = F32
    d |= 2u << 7;                       // a_format = TF32
    d |= 2u << 10;                      // b_format = TF32
    d |= (uint32_t)(BN >> 3) << 17;     // n_dim
    d |= (uint32_t)(BM >> 4) << 24;     // m_dim
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
}  // namespace tc

template <int BN>
__global__ void __launch_bounds__(tc::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
               const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, const tc::Sched sched,
               int M, int N, float alpha, float beta, float* __restrict__ C, int ldc, const float* __restrict__ bias) {
    using namespace tc;
    constexpr int STAGES = stages(BN), B_BYTES = b_bytes(BN), STAGE_BYTES = stage_bytes(BN);
    Span seg;
    if (!get_seg(sched, blockIdx.x, 0, seg)) return;       // nothing dealt to this CTA (uniform per CTA)

    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* scr_all = reinterpret_cast<float*>(base + (size_t)STAGES * STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)STAGES * STAGE_BYTES + (size_t)EPI_WARPS * SCR_FLOATS * 4);
    uint64_t* full = bars;                      // [STAGES]  TMA bytes landed
    uint64_t* empty = bars + STAGES;            // [STAGES]  MMAs that read the stage have retired
    uint64_t* tmem_full = bars + 2 * STAGES;    // [2]       accumulator complete
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;   // [2]   accumulator drained by the epilogue warps
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], EPI_WARPS); }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ---- TMA producer -----------------------------------------------------------------------------------
        if (lane == 0) {
            int it = 0;
            for (int i = 0; get_seg(sched, g, i, seg); i++) {
                const int m0 = seg.tm * BM, n0 = seg.tn * BN;
                for (int kb = seg.kb0; kb < seg.kb1; kb++, it++) {
                    const int s = it % STAGES, ph = it / STAGES;
                    if (ph > 0) mbar_wait_bounded(&empty[s], (ph - 1) & 1);
                    unsigned char* st = base + (size_t)s * STAGE_BYTES;
                    if (sched.dbg == 2) { mbar_arrive(&full[s]); continue; }
                    mbar_expect_tx(&full[s], STAGE_BYTES);
                    int ak = kb * BK, am = m0;
                    if (sched.tap_slabs > 0) {
                        const int tap = kb / sched.tap_slabs;
                        ak = (kb - tap * sched.tap_slabs) * BK;
                        am = m0 + sched.tap_row[tap];
                    }
                    if (sched.a_mn) {
                        tma_load_mn(st, &mapAh, m0, kb * BK, BM / 32, &full[s]);
                        tma_load_mn(st + A_BYTES, &mapAl, m0, kb * BK, BM / 32, &full[s]);
                    } else {
                        tma_load_2d(st, &mapAh, ak, am, &full[s]);
                        tma_load_2d(st + A_BYTES, &mapAl, ak, am, &full[s]);
                    }
                    if (sched.b_mn) {
                        // (b_col0 shifts along K = the ROW coordinate of an MN-major operand: any offset, no alignment constraint)
                        tma_load_mn(st + 2 * A_BYTES, &mapBh, n0, kb * BK + sched.b_col0, BN / 32, &full[s]);
                        tma_load_mn(st + 2 * A_BYTES + B_BYTES, &mapBl, n0, kb * BK + sched.b_col0, BN / 32, &full[s]);
                    } else {
                        tma_load_2d(st + 2 * A_BYTES, &mapBh, kb * BK + sched.b_col0, n0, &full[s]);
                        tma_load_2d(st + 2 * A_BYTES + B_BYTES, &mapBl, kb * BK + sched.b_col0, n0, &full[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issue --------------------------------------------------------------------------------------
        if (lane == 0) {
            const uint32_t idesc = make_idesc<BN>() | (sched.a_mn ? 1u << 15 : 0u) | (sched.b_mn ? 1u << 16 : 0u);
            // bytes (>> 4) from one K = 8 step to the next: 32 bytes inside the swizzle row (K-major) / one 8-row group (MN-major)
            const uint64_t a_step = sched.a_mn ? 1024 >> 4 : 32 >> 4, b_step = sched.b_mn ? 1024 >> 4 : 32 >> 4;
            int it = 0;
            for (int i = 0; get_seg(sched, g, i, seg); i++) {
                const int acc = i & 1;
                if (i >= 2) mbar_wait_bounded(&tmem_empty[acc], ((i >> 1) - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t td = tmem_base + (uint32_t)(acc * BN);
                for (int kb = seg.kb0; kb < seg.kb1; kb++, it++) {
                    const int s = it % STAGES, ph = it / STAGES;
                    mbar_wait_bounded(&full[s], ph & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_u32(base + (size_t)s * STAGE_BYTES);
                    const uint64_t dAh = sched.a_mn ? make_desc_mn(sa) : make_desc(sa);
                    const uint64_t dAl = sched.a_mn ? make_desc_mn(sa + A_BYTES) : make_desc(sa + A_BYTES);
                    const uint64_t dBh = sched.b_mn ? make_desc_mn(sa + 2 * A_BYTES) : make_desc(sa + 2 * A_BYTES);
                    const uint64_t dBl = sched.b_mn ? make_desc_mn(sa + 2 * A_BYTES + B_BYTES) : make_desc(sa + 2 * A_BYTES + B_BYTES);
                    if (sched.dbg != 1)
#pragma unroll
                    for (int k = 0; k < BK / 8; k++) {
                        const uint64_t ka = (uint64_t)k * a_step, kb_ = (uint64_t)k * b_step;
                        umma_tf32(td, dAl + ka, dBh + kb_, idesc, (kb != seg.kb0 || k != 0) ? 1u : 0u);
                        umma_tf32(td, dAh + ka, dBl + kb_, idesc, 1);
                        umma_tf32(td, dAh + ka, dBh + kb_, idesc, 1);
                    }
                    umma_commit(&empty[s]);                                           // stage reusable once these MMAs retire
                }
                umma_commit(&tmem_full[acc]);
            }
        }
    } else {
        // ---- epilogue (warps 2..9): warp e reads TMEM lane quarter (warp & 3), column half (e >> 2) ----------
        const int e = warp - 2;
        const int q = warp & 3;                                  // TMEM lane quarter this warp may access
        const int half = e >> 2;
        float* scr = scr_all + (size_t)e * SCR_FLOATS;
        const bool vec_ok = (ldc & 3) == 0 && (N & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
        for (int i = 0; get_seg(sched, g, i, seg); i++) {
            const int acc = i & 1;
            const int m0 = seg.tm * BM, n0 = seg.tn * BN;
            const bool direct = seg.kb0 == 0 && seg.kb1 == sched.nk;
            mbar_wait_bounded(&tmem_full[acc], (i >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
                if (n0 + c0 >= N) break;
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
#pragma unroll
                for (int j = 0; j < 32; j++) scr[lane * 33 + j] = __uint_as_float(r[j]);
                __syncwarp();
                if (vec_ok) {
                    // 4 rows x 8 float4 per pass: 16-byte stores / vector reductions (REDG.F32x4: a quarter of the L2 atomic ops)
                    const int rsub = lane >> 3, cq = (lane & 7) * 4;
                    const int n = n0 + c0 + cq;
                    if (n < N) {
                        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (bias && seg.kb0 == 0) bv = make_float4(bias[n], bias[n + 1], bias[n + 2], bias[n + 3]);
#pragma unroll
                        for (int rr = 0; rr < 32; rr += 4) {
                            const int row = m0 + q * 32 + rr + rsub;
                            if (row < M) {
                                const float* sp = scr + (rr + rsub) * 33 + cq;
                                float4 v = make_float4(alpha * sp[0] + bv.x, alpha * sp[1] + bv.y, alpha * sp[2] + bv.z, alpha * sp[3] + bv.w);
                                float* cp = C + (size_t)row * ldc + n;
                                if (direct) {
                                    if (beta != 0.f) {
                                        const float4 o = *reinterpret_cast<const float4*>(cp);
                                        v.x += beta * o.x; v.y += beta * o.y; v.z += beta * o.z; v.w += beta * o.w;
                                    }
                                    if (sched.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                                    *reinterpret_cast<float4*>(cp) = v;
                                } else {
                                    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(cp), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                                }
                            }
                        }
                    }
                } else {
                    const int n = n0 + c0 + lane;
                    if (n < N) {
                        const float bv = (bias && seg.kb0 == 0) ? bias[n] : 0.f;
#pragma unroll 4
                        for (int rr = 0; rr < 32; rr++) {
                            const int row = m0 + q * 32 + rr;
                            if (row < M) {
                                float* cp = C + (size_t)row * ldc + n;
                                float v = alpha * scr[rr * 33 + lane] + bv;
                                if (direct) {
                                    if (beta != 0.f) v += beta * (*cp);
                                    *cp = sched.relu ? fmaxf(v, 0.f) : v;
                                } else {
                                    atomicAdd(cp, v);
                                }
                            }
                        }
                    }
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
    }
}

// ---- operand preparation ---------------------------------------------------------------------------------------
// hi/lo split of a K-contiguous operand [rows, K] (pitch ld) into scratch with pitch Kp (multiple of 4);
// hi == nullptr: only lo is written (the tensor core ignores the low 13 mantissa bits of the raw operand).
__global__ void split_kernel(const float* __restrict__ src, int rows, int K, int ld, float* __restrict__ hi, float* __restrict__ lo, int Kp) {
    const int k4 = Kp >> 2;
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)rows * k4; idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / k4), k = (int)(idx - (long long)r * k4) * 4;
        const float* p = src + (size_t)r * ld + k;
        float4 v;
        if (vec && k + 3 < K) v = ldg_stream(p);
        else {
            v.x = k < K ? __ldg(p) : 0.f; v.y = k + 1 < K ? __ldg(p + 1) : 0.f;
            v.z = k + 2 < K ? __ldg(p + 2) : 0.f; v.w = k + 3 < K ? __ldg(p + 3) : 0.f;
        }
        float4 h;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
        const size_t o = (size_t)r * Kp + k;
        if (hi) *reinterpret_cast<float4*>(hi + o) = h;
        *reinterpret_cast<float4*>(lo + o) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
    }
}
// src is [K, rows] (pitch ld): transposed split into hi, lo [rows, Kp]
// shift: dst[r][k] = src[k + shift][r] (zeros beyond the end) -- TMA box origins must be 16-byte aligned, so an operand that
// has to be read at arbitrary K offsets (the weight gradient of a convolution tap) is prepared in four phases
__global__ void transpose_split_kernel(const float* __restrict__ src, int K, int rows, int ld, float* __restrict__ hi, float* __restrict__ lo, int Kp,
                                       int shift) {
    __shared__ float t[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x;          // source column = output row
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int k = blockIdx.y * 32 + i;
        t[i][threadIdx.x] = (k + shift < K && c < rows) ? __ldg(src + (size_t)(k + shift) * ld + c) : 0.f;
    }
    __syncthreads();
    const int k2 = blockIdx.y * 32 + threadIdx.x;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r2 = blockIdx.x * 32 + i;
        if (r2 < rows && k2 < Kp) {
            const float v = t[threadIdx.x][i];             // zero beyond K
            const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            hi[(size_t)r2 * Kp + k2] = h;
            lo[(size_t)r2 * Kp + k2] = v - h;
        }
    }
}
// zeroes the output tiles that take stream-K partial sums; blockIdx.y = one 16-row band of a tile
__global__ void zero_tiles_kernel(float* __restrict__ C, int ldc, int M, int N, int tiles_n, int first_tile, int BN) {
    const int tile = first_tile + blockIdx.x;
    const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
    const int r0 = tm * tc::BM + blockIdx.y * 16, c0 = tn * BN;
    const bool vec = (ldc & 3) == 0 && (N & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
    if (vec) {
        const int q = BN >> 2;
        for (int idx = threadIdx.x; idx < 16 * q; idx += blockDim.x) {
            const int r = r0 + idx / q, c = c0 + (idx % q) * 4;
            if (r < M && c < N) *reinterpret_cast<float4*>(C + (size_t)r * ldc + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        for (int idx = threadIdx.x; idx < 16 * BN; idx += blockDim.x) {
            const int r = r0 + idx / BN, c = c0 + idx % BN;
            if (r < M && c < N) C[(size_t)r * ldc + c] = 0.f;
        }
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}
// [rows, K] fp32, K contiguous, row pitch ld floats; box = 32 floats x box_rows, 128-byte swizzle, OOB -> 0
static bool make_map(CUtensorMap* map, const float* ptr, int rows, int K, int ld, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)tc::BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// MN-major operand: [K, rows] fp32 with `rows` contiguous (pitch ld floats): box = 32 columns x 32 K-rows (one chunk of Sched::a_mn)
static bool make_map_mn(CUtensorMap* map, const float* ptr, int rows, int K, int ld) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)K};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)tc::BK};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
static int tc_enabled() {
    static int v = -1;
    if (v < 0) v = env_int("S2S_TC", 1);
    return v;
}
// S2S_TC_RAWHI=1 (default): a K-contiguous, 16-byte-aligned operand is used in place as the hi part (the tf32 MMA reads
// only the top 19 bits); 0 = always materialise the masked hi copy.
static int tc_rawhi() {
    static int v = -1;
    if (v < 0) v = env_int("S2S_TC_RAWHI", 1);
    return v;
}

// S2S_TC_MN=1 (default): operands of the NN / TN forms whose MN extent and pitch are multiples of 4 floats (16-byte aligned) are consumed in place as
// MN-major UMMA operands -- no transposing pre-pass, only the low part is written (in the operand's own layout); 0 = always transpose.
static int tc_mn() {
    static int v = -1;
    if (v < 0) v = env_int("S2S_TC_MN", 1);
    return v;
}

struct TcOp { const float* hi = nullptr; const float* lo = nullptr; int ld_hi = 0, ld_lo = 0; bool mn = false; };

// src: K-contiguous [rows, K] (transposed == false) or [K, rows] (transposed == true), pitch ld
static int tc_prepare(s2s_ctx* ctx, const float* src, int rows, int K, int ld, bool transposed, TcOp* op) {
    const int Kp = (K + 3) & ~3;
    op->mn = false;
    if (ctx->tc_cache_on) {
        for (const TcCacheEntry& c : ctx->tc_cache) {
            if (c.transposed != transposed || c.K != K || c.ld != ld) continue;
            const ptrdiff_t off = src - c.src;
            if (c.mn) {     // cached in place ([K, cols], low part with pitch c.ld_lo): any 16-byte-aligned block of columns
                if (off >= 0 && off + rows <= c.cols && (off & 3) == 0) {
                    op->hi = src; op->lo = c.lo + off; op->ld_hi = ld; op->ld_lo = c.ld_lo; op->mn = true;
                    return 0;
                }
                continue;
            }
            // transposed: a block of `rows` source columns starting `off` columns into the cached matrix
            if (transposed ? (off >= 0 && off + rows <= c.cols) : (off == 0 && rows <= c.cols)) {
                const size_t r0 = transposed ? (size_t)off : 0;
                op->hi = c.hi + r0 * c.ld_hi; op->lo = c.lo + r0 * c.ld_lo; op->ld_hi = c.ld_hi; op->ld_lo = c.ld_lo;
                return 0;
            }
        }
    }
    float *hi = nullptr, *lo;
    if (transposed && tc_mn() && (rows & 3) == 0 && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        // MN-major in place: hi = the raw operand (the tf32 MMA reads only the top 19 bits), lo in the same [K, rows] layout
        S2S_ALLOC(lo, ctx->arena, float, (size_t)K * rows);
        const long long n4 = (long long)K * (rows >> 2);
        int blocks = (int)((n4 + 255) / 256);
        if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
        split_kernel<<<blocks, 256, 0, ctx->stream>>>(src, K, rows, ld, nullptr, lo, rows);
        S2S_LAUNCH_CHECK(ctx);
        op->hi = src; op->ld_hi = ld; op->lo = lo; op->ld_lo = rows; op->mn = true;
        if (ctx->tc_cache_on) ctx->tc_cache.push_back(TcCacheEntry{src, K, rows, ld, transposed, op->hi, op->lo, op->ld_hi, op->ld_lo, true});
        return 0;
    }
    S2S_ALLOC(lo, ctx->arena, float, (size_t)rows * Kp);
    const bool raw = !transposed && tc_rawhi() && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    if (!raw) S2S_ALLOC(hi, ctx->arena, float, (size_t)rows * Kp);
    if (transposed) {
        transpose_split_kernel<<<dim3(ceil_div(rows, 32), ceil_div(Kp, 32)), dim3(32, 8), 0, ctx->stream>>>(src, K, rows, ld, hi, lo, Kp, 0);
    } else {
        const long long n4 = (long long)rows * (Kp >> 2);
        int blocks = (int)((n4 + 255) / 256);
        if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
        split_kernel<<<blocks, 256, 0, ctx->stream>>>(src, rows, K, ld, hi, lo, Kp);
    }
    S2S_LAUNCH_CHECK(ctx);
    op->hi = raw ? src : hi; op->ld_hi = raw ? ld : Kp;
    op->lo = lo; op->ld_lo = Kp;
    if (ctx->tc_cache_on) ctx->tc_cache.push_back(TcCacheEntry{src, K, rows, ld, transposed, op->hi, op->lo, op->ld_hi, op->ld_lo, false});
    return 0;
}

struct TcConv { bool relu = false; int tap_slabs = 0; int tap_row[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; int b_col0 = 0; int a_cols = 0; };

// C[M,N] = alpha A' B^T (+beta C)(+bias) with prepared (K-contiguous, hi/lo) operands.  Plain GEMM: A' = A [M,K].  With
// cv.tap_slabs > 0 the A operand is the [M, a_cols] activation matrix read at 9 row offsets (implicit 3x3 convolution).
static int tc_run(s2s_ctx* ctx, int M, int N, int K, float alpha, const TcOp& a, const TcOp& b, float beta, float* C, int ldc, const float* bias,
                  const TcConv& cv) {
    using namespace tc;
    const int BN = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
    Sched sc;
    sc.tiles_m = ceil_div(M, BM); sc.tiles_n = ceil_div(N, BN); sc.nk = ceil_div(K, BK);
    sc.G = ctx->gemm_sm_limit > 0 ? std::min(ctx->sm_count, ctx->gemm_sm_limit) : ctx->sm_count;
    static const int dbg_mode = env_int("S2S_TC_DBG", 0), tail_on = env_int("S2S_TC_TAIL", 1);
    sc.dbg = dbg_mode;
    sc.tap_slabs = cv.tap_slabs; sc.b_col0 = cv.b_col0; sc.relu = cv.relu ? 1 : 0;
    sc.a_mn = a.mn ? 1 : 0; sc.b_mn = b.mn ? 1 : 0;
    S2S_REQUIRE(!(a.mn && cv.tap_slabs > 0), "gemm_tc: the tap-shifted A operand of the implicit convolution is K-major");
    for (int t = 0; t < 9; t++) sc.tap_row[t] = cv.tap_row[t];
    const int Tt = sc.tiles_m * sc.tiles_n;
    sc.R = Tt / sc.G; sc.rem = Tt - sc.R * sc.G;
    // tail: cut the left-over tiles along K into >= MIN_SHARE-slab shares; if that gives no more CTAs than tiles, or C
    // cannot take atomic partial sums (beta other than 0 / 1), the tail is an ordinary partial round of whole tiles
    const long long U = (long long)sc.rem * sc.nk;
    sc.Gp = (int)(U / MIN_SHARE < sc.G ? U / MIN_SHARE : sc.G);
    if (sc.Gp <= sc.rem || (beta != 0.f && beta != 1.f) || !tail_on || cv.relu) sc.Gp = sc.rem;
    const bool atomics = sc.Gp > sc.rem;

    const int aK = cv.tap_slabs > 0 ? cv.a_cols : K;           // extent of the A operand's K axis in memory
    const int bK = K;                                          // true extent: shifted slabs beyond it read zeros
    CUtensorMap mAh, mAl, mBh, mBl;
    const bool okA = a.mn ? make_map_mn(&mAh, a.hi, M, aK, a.ld_hi) && make_map_mn(&mAl, a.lo, M, aK, a.ld_lo)
                          : make_map(&mAh, a.hi, M, aK, a.ld_hi, BM) && make_map(&mAl, a.lo, M, aK, a.ld_lo, BM);
    const bool okB = b.mn ? make_map_mn(&mBh, b.hi, N, bK, b.ld_hi) && make_map_mn(&mBl, b.lo, N, bK, b.ld_lo)
                          : make_map(&mBh, b.hi, N, bK, b.ld_hi, BN) && make_map(&mBl, b.lo, N, bK, b.ld_lo, BN);
    if (!okA || !okB)
        return fail("gemm_tc: cuTensorMapEncodeTiled failed (M=%d N=%d K=%d)", M, N, K);
    static bool attr = false;
    if (!attr) {
        S2S_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(64)));
        S2S_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(128)));
        S2S_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(256)));
        attr = true;
    }
    if (atomics && beta == 0.f) {
        zero_tiles_kernel<<<dim3(sc.rem, BM / 16), 256, 0, ctx->stream>>>(C, ldc, M, N, sc.tiles_n, sc.R * sc.G, BN);
        S2S_LAUNCH_CHECK(ctx);
    }
    if (dbg_mode == 9) {
        fprintf(stderr, "[gemm_tc] M=%d N=%d K=%d BN=%d tiles=%dx%d nk=%d R=%d rem=%d Gp=%d taps=%d row0=%d row8=%d bcol=%d beta=%g\n", M, N, K, BN, sc.tiles_m,
                sc.tiles_n, sc.nk, sc.R, sc.rem, sc.Gp, sc.tap_slabs, sc.tap_row[0], sc.tap_row[8], sc.b_col0, (double)beta);
        cudaStreamSynchronize(ctx->stream);
    }
    prof_begin(ctx, S2S_PROF_GEMM);
    if (BN == 64) gemm_tc_kernel<64><<<sc.G, THREADS, smem_bytes(64), ctx->stream>>>(mAh, mAl, mBh, mBl, sc, M, N, alpha, beta, C, ldc, bias);
    else if (BN == 128) gemm_tc_kernel<128><<<sc.G, THREADS, smem_bytes(128), ctx->stream>>>(mAh, mAl, mBh, mBl, sc, M, N, alpha, beta, C, ldc, bias);
    else gemm_tc_kernel<256><<<sc.G, THREADS, smem_bytes(256), ctx->stream>>>(mAh, mAl, mBh, mBl, sc, M, N, alpha, beta, C, ldc, bias);
    prof_end(ctx, S2S_PROF_GEMM, 2.0 * M * N * (double)K);
    ctx->kcount[S2S_KC_GEMM_TC]++;
    S2S_LAUNCH_CHECK(ctx);
    if (dbg_mode == 9) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        fprintf(stderr, "[gemm_tc]   -> %s\n", cudaGetErrorString(e));
    }
    return 0;
}

// ---- implicit 3x3 convolution on channels-last activations kept on their FULL grid [nb, Hh, Ww, C] -------------------------
// (rows = pixels; the valid region shrinks by 2 per convolution, rows outside it hold don't-care values)
//   forward        out[m, n]  = sum_t sum_c in[m + off_t, c] Wp[n, t*C + c] + bias[n]          off_t = kh*Ww + kw
//   data gradient  din[m, c]  = sum_t sum_n dout[m - off_t, n] WpT[c, t*N + n]                  (dout zero outside its valid region)
//   weight grad.   dWp[n, t*C + c] += sum_m dout[m, n] in[m + off_t, c]                         one launch per tap, K = pixels
// C and N must be multiples of 32 (one tap = whole 128-byte slabs).
int conv3_tc_forward(s2s_ctx* ctx, const float* in, int64_t Mg, int Ww, int C, const float* Wp, const float* bias, int N, float* out, bool relu) {
    S2S_REQUIRE(C % 32 == 0 && Mg < (1ll << 31), "conv3_tc_forward: C must be a multiple of 32");
    if (!get_encode()) return fail("conv3_tc: cuTensorMapEncodeTiled unavailable");
    TcOp a, b;
    S2S_TRY(tc_prepare(ctx, in, (int)Mg, C, C, false, &a));
    S2S_TRY(tc_prepare(ctx, Wp, N, 9 * C, 9 * C, false, &b));
    TcConv cv; cv.tap_slabs = C / 32; cv.a_cols = C; cv.relu = relu;
    for (int t = 0; t < 9; t++) cv.tap_row[t] = (t / 3) * Ww + (t % 3);
    return tc_run(ctx, (int)Mg, N, 9 * C, 1.f, a, b, 0.f, out, N, bias, cv);
}
// WpT [C, 9N]: WpT[c, t*N + n] = Wp[n, t*C + c] (prepared by the caller)
int conv3_tc_dgrad(s2s_ctx* ctx, const float* dout, int64_t Mg, int Ww, int N, const float* WpT, int C, float* din) {
    S2S_REQUIRE(N % 32 == 0 && Mg < (1ll << 31), "conv3_tc_dgrad: N must be a multiple of 32");
    TcOp a, b;
    S2S_TRY(tc_prepare(ctx, dout, (int)Mg, N, N, false, &a));
    S2S_TRY(tc_prepare(ctx, WpT, C, 9 * N, 9 * N, false, &b));
    TcConv cv; cv.tap_slabs = N / 32; cv.a_cols = N;
    for (int t = 0; t < 9; t++) cv.tap_row[t] = -((t / 3) * Ww + (t % 3));
    return tc_run(ctx, (int)Mg, C, 9 * N, 1.f, a, b, 0.f, din, C, nullptr, cv);
}
int conv3_tc_wgrad(s2s_ctx* ctx, const float* dout, const float* in, int64_t Mg, int Ww, int N, int C, float* dWp) {
    S2S_REQUIRE(Mg < (1ll << 31), "conv3_tc_wgrad: too many pixels");
    TcOp a, bm;
    S2S_TRY(tc_prepare(ctx, dout, N, (int)Mg, N, true, &a));          // dout^T [N, Mg]
    const int K = (int)Mg, Kp = (K + 3) & ~3;
    S2S_TRY(tc_prepare(ctx, in, C, K, C, true, &bm));                 // in^T [C, Mg]
    if (bm.mn) {
        // consumed in place as an MN-major operand: a tap is a shift along K = the row coordinate of the TMA box, so the nine products
        // share one prepared operand
        for (int t = 0; t < 9; t++) {
            TcConv cv; cv.b_col0 = (t / 3) * Ww + (t % 3);
            S2S_TRY(tc_run(ctx, N, C, K, 1.f, a, bm, 1.f, dWp + (size_t)t * C, 9 * C, nullptr, cv));
        }
        return 0;
    }
    // K-major copies: TMA box origins must be 16-byte aligned along the contiguous axis, so in^T is prepared shifted by 0..3 pixels:
    // tap offset = 4 q + phase, the aligned part goes into the TMA coordinate
    TcOp b[4];
    b[0] = bm;
    for (int ph = 1; ph < 4; ph++) {
        float *hi, *lo;
        S2S_ALLOC(hi, ctx->arena, float, (size_t)C * Kp);
        S2S_ALLOC(lo, ctx->arena, float, (size_t)C * Kp);
        transpose_split_kernel<<<dim3(ceil_div(C, 32), ceil_div(Kp, 32)), dim3(32, 8), 0, ctx->stream>>>(in, K, C, C, hi, lo, Kp, ph);
        S2S_LAUNCH_CHECK(ctx);
        b[ph].hi = hi; b[ph].lo = lo; b[ph].ld_hi = Kp; b[ph].ld_lo = Kp;
    }
    for (int t = 0; t < 9; t++) {
        const int off = (t / 3) * Ww + (t % 3);
        TcConv cv; cv.b_col0 = off & ~3;
        S2S_TRY(tc_run(ctx, N, C, K, 1.f, a, b[off & 3], 1.f, dWp + (size_t)t * C, 9 * C, nullptr, cv));
    }
    return 0;
}

// C = alpha op(A) op(B) (+beta C)(+bias) on the tensor cores for the three operand orders the model uses:
//   NT  C = A B^T      (forward projections)                 both operands already K-contiguous
//   NN  C = A B        (data gradients: dX = dA W_x, dh = dVh W_V)      B transposed by the split pre-pass
//   TN  C += A^T B     (weight gradients, K = B*L)                      both transposed by the split pre-pass
// Small products stay on the exact-fp32 SIMT kernel (launch-bound anyway); `force` (test hook) lifts the size threshold.
int gemm_tc_f32(s2s_ctx* ctx, bool tA, bool tB, int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb,
                float beta, float* C, int ldc, const float* bias, bool* handled, bool force, bool relu) {
    using namespace tc;
    *handled = false;
    if (!tc_enabled() && !force) return 0;
    if (K < 1 || M < 1 || N < 1) return 0;
    if (tA && tB) return 0;
    if (!force && (double)M * N * K < 5e8) return 0;
    if (!get_encode()) return 0;

    TcOp a, b;
    S2S_TRY(tc_prepare(ctx, A, M, K, lda, tA, &a));                       // tA: A is [K, M]
    S2S_TRY(tc_prepare(ctx, B, N, K, ldb, !tB, &b));                      // !tB: B is [K, N]
    TcConv plain; plain.relu = relu;
    S2S_TRY(tc_run(ctx, M, N, K, alpha, a, b, beta, C, ldc, bias, plain));
    *handled = true;
    return 0;
}

}  // namespace s2s

// ---- test hooks for the implicit convolution (flattened-grid semantics, see above) ------------------------------------
using namespace s2s;
extern "C" {
__attribute__((visibility("default"))) int s2s_conv3_forward(s2s_ctx* ctx, const float* in, int64_t Mg, int Ww, int C, const float* Wp, const float* bias,
                                                             int N, float* out, int relu) {
    S2S_REQUIRE(ctx && in && Wp && out, "conv3_forward: null argument");
    ctx->arena.reset();
    return conv3_tc_forward(ctx, in, Mg, Ww, C, Wp, bias, N, out, relu != 0);
}
__attribute__((visibility("default"))) int s2s_conv3_dgrad(s2s_ctx* ctx, const float* dout, int64_t Mg, int Ww, int N, const float* WpT, int C, float* din) {
    S2S_REQUIRE(ctx && dout && WpT && din, "conv3_dgrad: null argument");
    ctx->arena.reset();
    return conv3_tc_dgrad(ctx, dout, Mg, Ww, N, WpT, C, din);
}
__attribute__((visibility("default"))) int s2s_conv3_wgrad(s2s_ctx* ctx, const float* dout, const float* in, int64_t Mg, int Ww, int N, int C, float* dWp) {
    S2S_REQUIRE(ctx && dout && in && dWp, "conv3_wgrad: null argument");
    ctx->arena.reset();
    return conv3_tc_wgrad(ctx, dout, in, Mg, Ww, N, C, dWp);
}
}
