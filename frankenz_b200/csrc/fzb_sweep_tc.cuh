// Tensor-core sweep (tcgen05 + TMEM) for the free-scale / no-model-error likelihood (FS0), included by fzb_fast.cu.
//
// Per pair the packed FP32 sweep (k_sweep2) spends 15 of its ~35 FMA-pipe operations on three dot products with the
// model fluxes: inter = sum (w d)_b m_b, shape = sum w_b m_b^2 (-> optimal scale, pdf.py:181-185) and the first-order
// correction for the low half of the float64 data, G = sum (2 w d_lo)_b m_b.  They are GEMMs with K = Nf, so here the
// 5th-generation tensor cores compute them:  D[object][model] = A[object][K] . B[model][K]^T,  kind::tf32, M = 128,
// N = 32, operands K-major in shared memory without swizzle, accumulators in TMEM, one elected thread issuing.
// tf32 keeps 11 significant bits, so inter and shape use the split x = hi + lo on both sides and three products per
// band (hi*hi + hi*lo + lo*hi: 3 Nf K-slots, two K = 8 instructions up to five bands, three for six), which leaves a relative error of ~5e-7 in
// the scale; the scale is the minimiser of chi2, so that error enters chi2 only as S/N^2 (ds/s)^2 (envelope theorem),
// ~1e-12 S/N^2.  G takes the same split: the correction K_o - s G = sum g_b (d_b - s m_b) is small, but K_o and s G are
// each ~1e-7 S/N^2 and cancel, so G needs fp32-level accuracy as well.  The residual part,
// chi2 = K_o - s G + sum_b w_b (d_b - s m_b)^2, the ln-likelihood and the online reductions stay in packed FP32 on the
// CUDA cores exactly as in k_sweep2, now 20 FMA-pipe operations per pair.
//
// Sizing.  tools/tc_rate.cu: one tcgen05.mma of this kind (M = 128, K = 8, operands in shared memory) takes ~118
// cycles whatever N is (16 .. 64), so the six MMAs of an (M-tile, chunk) must cover enough pairs: N = 32 gives
// 128 x 32 pairs per 708 cycles = 5.8 pairs / clk / SM (1.7e12 pairs/s per GPU), above the MUFU / FMA ceilings of the
// CUDA-core part.  TMEM: 2 M-tiles x 3 products x 32 columns, double buffered = 384 of 512 columns.
//
// CTA = 18 warps, 256 objects: warps 0-15 are four consumer warpgroups (warp w reads TMEM lanes 32 (w % 4) ..); warpgroup
// g works on M-tile g >> 1 and, of every 32-model chunk, on the two 8-model sub-batches of parity g & 1, so an object
// is shared by two threads that each see half of the models (their partial reductions are merged by k_merge like two
// model splits; pass 2 accumulates with atomics anyway).  Every consumer thread evaluates two models at a time in the
// two halves of packed registers.  Warp 16 stages 256-model tiles with TMA bulk copies, warp 17 issues the MMAs.
//
// Shared-memory / global layout of a model tile (TC_TILE_BYTES, one bulk copy):
//   [k-step s (4)][row group (32)][k chunk (2)][row (8)][4 floats]   MMA B operand, SBO = 256 B, LBO = 128 B
//        k-steps 0-1: m_hi | m_lo | m_hi | 0        (paired with  -x_hi | -x_hi | -x_lo | 0,  x = w d, for inter
//                                                    and with      g_hi |  g_hi |  g_lo | 0,  g = 2 w d_lo, for G)
//        k-steps 2-3: q_hi | q_lo | q_hi | 0        (q = m^2;     w_hi |  w_hi |  w_lo | 0)
//   [model pair (128)][6] float2   (m_even, m_odd) per band, then the prior pair   (48 B, three LDS.128); with MLO
//        (models that are not fp32-representable, e.g. the float64 grid of simulate.make_model_grid) the float64
//        remainders (m - float(m))_even/odd per band sit between the bands and the prior (nf more float2)
//   [model pair (128)]    {invnorm_even, invnorm_odd, bin_even, bin_odd}
//   [8 models (32)]       {bin of the first, 1 if all eight share it, 1/norm of that bin (same bin = same kernel
//                          width and grid position = same edge normalisation), -}
#pragma once

constexpr int TC_TM = 256;                       // models per shared-memory tile
constexpr int TC_NC = 32;                        // models per MMA / TMEM chunk
constexpr int TC_NSUB = TC_NC / 8;               // 8-model sub-batches per chunk
constexpr int TC_NWG = 4;                        // consumer warpgroups
constexpr int TC_MT = 2;                         // 128-object M-tiles per CTA
constexpr int TC_SPLIT = TC_NWG / TC_MT;         // warpgroups (threads) sharing an object = partial results per object
constexpr int TC_OBJS = TC_MT * 128;             // objects per CTA
constexpr int TC_THREADS = TC_NWG * 128 + 64;    // + TMA warp + MMA warp
constexpr int TC_CW = TC_NWG * 4;                // consumer warps; warp TC_CW = TMA, TC_CW + 1 = MMA
// K layout: the three products of a band (hi*hi, hi*lo, lo*hi) take the K-slots b, S + b, 2 S + b (S = 5 up to five
// bands, 6 for six), i.e. KS = 2 (3 for six bands) K = 8 instructions per product
__host__ __device__ constexpr int tc_slot(int nf) { return nf <= 5 ? 5 : nf; }
__host__ __device__ constexpr int tc_ks(int nf) { return (3 * tc_slot(nf) + 7) / 8; }
__host__ __device__ constexpr int tc_ksteps(int nf) { return 2 * tc_ks(nf); }      // model operand: m rows, m^2 rows
__host__ __device__ constexpr int tc_aksteps(int nf) { return 3 * tc_ks(nf); }     // object operand: -x, w, g
// 16-byte words per model pair: nf bands (+ nf bands of the float64 remainder, MLO) + prior
__host__ __device__ constexpr int tc_pairq(int nf, bool mlo = false) { return ((mlo ? 2 * nf : nf) + 2) / 2; }
__host__ __device__ constexpr int tc_opsec(int nf) { return tc_ksteps(nf) * TC_TM * 32; }
__host__ __device__ constexpr int tc_pairsec(int nf, bool mlo = false) { return (TC_TM / 2) * 16 * tc_pairq(nf, mlo); }
constexpr int TC_TAILSEC = (TC_TM / 2) * 16;
constexpr int TC_SUBSEC = (TC_TM / 8) * 16;        // per 8 models: {KDE bin of the first, 1 if all eight share it, 1/norm of that bin}
__host__ __device__ constexpr int tc_tile_bytes(int nf, bool mlo = false) { return tc_opsec(nf) + tc_pairsec(nf, mlo) + TC_TAILSEC + TC_SUBSEC; }
__host__ __device__ constexpr int tc_obja_tile(int nf) { return tc_aksteps(nf) * 128 * 32; }
__host__ __device__ constexpr int tc_obja_bytes(int nf) { return TC_MT * tc_obja_tile(nf); }
constexpr int TC_NSTAGE = 2;
constexpr int TC_CHUNK_COLS = TC_MT * 3 * TC_NC;  // TMEM columns of one chunk buffer
constexpr int TC_TMEM_COLS = 512;
__host__ __device__ constexpr size_t tc_smem(int nf, bool mlo = false) { return (size_t)tc_obja_bytes(nf) + (size_t)TC_NSTAGE * tc_tile_bytes(nf, mlo) + 512; }
static_assert(tc_tile_bytes(5) == 41472 && tc_tile_bytes(5) % 128 == 0 && tc_tile_bytes(6) % 128 == 0, "tile layout");
static_assert(tc_tile_bytes(4, true) % 128 == 0 && tc_tile_bytes(5, true) % 128 == 0 && tc_tile_bytes(6, true) % 128 == 0, "tile layout");
static_assert(tc_smem(6, true) <= 227 * 1024, "shared memory budget");
static_assert(2 * TC_CHUNK_COLS <= TC_TMEM_COLS, "TMEM budget");
static_assert(TC_SPLIT == 2 && TC_NSUB == 4, "sub-batch assignment below assumes two warpgroups per M-tile, four sub-batches");

__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// mbarrier wait that lets the hardware suspend the warp until the phase completes (suspend-time hint) instead of
// polling: a polling producer warp takes issue slots from the consumer warps that share its scheduler
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITS_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONES_%=;\n"
        "bra WAITS_%=;\n"
        "DONES_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// shared-memory matrix descriptor: no swizzle, K-major; core matrices (8 rows x 16 B) LBO apart along K, SBO apart
// along the rows (verified on the hardware by tools/tc_probe.cu)
__device__ __forceinline__ uint64_t tc_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// D[tmem] (+)= A[smem] . B[smem]^T; the descriptors are passed as (low word, high word): the high word (SBO, version) is
// the same for every operand, the low word is base + (byte offset >> 4)
template <bool ACC>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc) {
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 da, db;\n"
        "mov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\n"
        "setp.ne.b32 p, %5, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "n"(ACC ? 1 : 0)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
#define TC_TIE8(v) "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7])

// chi2 part of one object against a pair of models: sum_b w_b (d_b - s m_b)^2 + K - s G in the units of the weights
// MLO: m = float(m) + m_lo; the residual d - s m is formed with both parts (the second FMA adds -s m_lo to the already
// cancelled difference, so nothing of the remainder is lost to rounding; the scale itself needs no correction,
// envelope theorem)
template <int NF, bool MLO>
__device__ __forceinline__ f2 tc_pair_c(f2 B2, f2 C2, f2 G2, const f2* __restrict__ m2, const f2* d2, const f2* w2, f2 K2) {
    const f2 rc = pack2(fast_rcp(lo2(C2)), fast_rcp(hi2(C2)));
    const f2 ns = mul2(B2, rc);          // minus the optimal scale (the A operand of the inter product is negated)
    f2 c = fma2(ns, G2, K2);             // first-order correction for the low half of the data
#pragma unroll
    for (int b = 0; b < NF; ++b) {
        f2 r = fma2(ns, m2[b], d2[b]);
        if (MLO) r = fma2(ns, m2[NF + b], r);
        f2 t = mul2(r, w2[b]);
        c = fma2(t, r, c);
    }
    return c;
}
// ln-likelihood (log2 units, up to the per-object constant); the weights carry -log2(e)/2: c = -chi2 log2(e)/2
template <int NF, bool DP, bool PRIOR, bool TAIL, bool MLO>
__device__ __forceinline__ f2 tc_pair_l(f2 B2, f2 C2, f2 G2, const f2* __restrict__ m2, f2 prior2, const f2* d2, const f2* w2,
                                        f2 K2, f2 A2) {
    f2 c = tc_pair_c<NF, MLO>(B2, C2, G2, m2, d2, w2, K2);
    const f2 cc = c;
    if (PRIOR) c = add2(c, prior2);
    f2 l = c;
    if (DP) l = fma2(A2, pack2(fast_lg2(fabsf(lo2(cc))), fast_lg2(fabsf(hi2(cc)))), c);
    if (TAIL) l = pack2(lo2(l), -FLT_MAX);   // odd model count: the second model of the last pair is padding
    return l;
}
// exact 2^k for an integer-valued float k (0 below the normal range)
__device__ __forceinline__ float pow2i(float k) {
    return (k >= -126.f) ? __int_as_float(((int)fminf(k, 127.f) + 127) << 23) : 0.f;
}

// FUSE (pass 1 of the linear-domain form only): the single-pass variant.  A coarse pre-pass (every 16th model, the same
// kernel) has left a lower bound M0 of the object's maximum; the frame of the weights is fixed to R = -floor(M0) and
// every weight above the RUNNING cut wt_thresh max(2^(M0 + R), largest weight so far) goes into the KDE histogram right
// away.  The running cut never exceeds the final one, so the histogram holds a superset of the selection, and the
// surplus - weights between the cut of their moment and the final cut - lies within a factor fz_gfac above the former
// as long as the final maximum stays within that factor of 2^M0 (k_merge checks it).  Those weights (and the ones
// within the fp32 error below the cut) are recorded per thread and re-decided in float64 by k_fuse_fix.  Objects that
// break the assumptions (frame change, record overflow, maximum far above M0: the bright ones, whose posteriors are
// narrow) get their histogram row cleared and take the pruned pass 2 as before.
// FUSE = 2 is the same sweep without the histogram: M0 only seeds the frame and the live bits of pass 2, which then are
// nearly the final selection instead of the superset under the running maximum - the variant for the bright objects.
template <int NF, bool DP, bool PRIOR, int PASS, bool LIN = false, bool MLO = false, int FUSE = 0>
__global__ void __launch_bounds__(TC_THREADS, 1) k_sweep_tc(SweepParams P, const unsigned char* __restrict__ tiles, uint32_t lbo,
                                                            uint32_t sbo) {
    static_assert(NF >= 1 && NF <= 6, "filters per object");
    static_assert(!FUSE || (LIN && PASS == 1), "the fused variant is pass 1 of the linear-domain form");
    constexpr int SLOT = tc_slot(NF), KS = tc_ks(NF), AKSTEPS = tc_aksteps(NF), PQ = tc_pairq(NF, MLO);
    constexpr int TILE_BYTES = tc_tile_bytes(NF, MLO), OPSEC = tc_opsec(NF), PAIRSEC = tc_pairsec(NF, MLO);
    constexpr int PRI = MLO ? 2 * NF : NF;      // position of the prior pair among the float2 of a model pair
    constexpr int OBJA_TILE = tc_obja_tile(NF), OBJA_BYTES = tc_obja_bytes(NF);
    static_assert(!LIN || DP, "the linear-domain form is the dim_prior likelihood with (dof/2 - 1) = 1");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* objA = smem_raw;
    unsigned char* stage = smem_raw + OBJA_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage + (size_t)TC_NSTAGE * TILE_BYTES);
    uint64_t* tile_full = bars;                       // [NSTAGE]  TMA -> MMA + consumers
    uint64_t* tile_empty = bars + 2;                  // [NSTAGE]  MMA commit + consumer warps -> TMA
    uint64_t* acc_full = bars + 4;                    // [mt][buf] MMA commit -> consumers
    uint64_t* acc_empty = bars + 4 + 2 * TC_MT;       // [mt][buf] consumer warps of the M-tile -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 + 4 * TC_MT);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    const int64_t ntiles_all = (P.nm + TC_TM - 1) / TC_TM;
    const int64_t t0 = (int64_t)blockIdx.y * P.tiles_per_split;
    int64_t t1 = t0 + P.tiles_per_split;
    if (t1 > ntiles_all) t1 = ntiles_all;
    const int nt = (int)(t1 - t0);
    const int64_t tile_base = (int64_t)blockIdx.x * TC_OBJS;

    if (tid == 0) {
        for (int s = 0; s < TC_NSTAGE; ++s) { mbar_init(&tile_full[s], 1); mbar_init(&tile_empty[s], 1 + TC_CW); }
        for (int i = 0; i < 2 * TC_MT; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4 * TC_SPLIT); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_CW + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }

    // ---- consumer state --------------------------------------------------------------------------------------
    const int wg = warp >> 2;                   // consumer warpgroup
    const int mt = wg >> 1;                     // its M-tile
    const int half = wg & 1;                    // its sub-batch parity
    const int row = (warp & 3) * 32 + lane;     // row of the M-tile = TMEM lane
    f2 d2[NF], w2[NF], K2 = 0, A2 = 0;
    f2 M2 = 0, S2 = 0, acc2 = 0;
    float thr = 0.f, Mfl = -FLT_MAX;
    float Rf = 0.f;                             // LIN: reference exponent of the linear-domain weights
    double Sd = 0.0;
    int best0 = 0, best1 = 0, oidx = -1;
    float Yc = 0.f;                             // FUSE: running maximum of the weights behind the cut (seeded by M0)
    float fcut = FLT_MAX;                       // FUSE: the running cut wt_thresh Yc (FLT_MAX: object not fused)
    bool fuse_ok = false;                       // FUSE: frame still the one fixed by M0
    int rcnt = 0, fz_seg = 0;                   // FUSE: records written by this thread, its record segment
    if (warp < TC_CW) {
        const int64_t slot = tile_base + (int64_t)mt * 128 + row;
        int64_t o;
        if (PASS == 1 && !P.objlist) o = slot < P.No_pad ? slot : P.No_pad - 1;
        else o = slot < P.No ? P.objlist[slot] : -1;      // pass 2, or pass 1 over a list of objects
        oidx = (int)o;
        const int64_t oo = o < 0 ? 0 : o;
        float arow[8 * AKSTEPS];
#pragma unroll
        for (int i = 0; i < 8 * AKSTEPS; ++i) arow[i] = 0.f;
        float kk = 0.f;
#pragma unroll
        for (int b = 0; b < NF; ++b) {
            const float d = P.od[b * P.No_pad + oo], w = P.ow[b * P.No_pad + oo];
            const float x = P.ox[b * P.No_pad + oo], dl = P.odl[b * P.No_pad + oo];
            d2[b] = pack2(d, d);
            w2[b] = LIN ? pack2(-w, -w) : pack2(w, w);      // LIN: positive weights, x = chi2 log2(e)/2 >= 0
            const float g = LIN ? -(w * dl) : w * dl;
            kk = fmaf(g, d, kk);
            const float xh = tf32_rn(x), xl = tf32_rn(x - xh);
            const float wh = tf32_rn(w), wl = tf32_rn(w - wh);
            arow[b] = -xh; arow[SLOT + b] = -xh; arow[2 * SLOT + b] = -xl;
            arow[8 * KS + b] = wh; arow[8 * KS + SLOT + b] = wh; arow[8 * KS + 2 * SLOT + b] = wl;
            const float gh = tf32_rn(g), gl = tf32_rn(g - gh);
            arow[16 * KS + b] = gh; arow[16 * KS + SLOT + b] = gh; arow[16 * KS + 2 * SLOT + b] = gl;
        }
        K2 = pack2(kk, kk);
        const float a = P.oA[oo];
        A2 = pack2(a, a);
        if (half == 0) {        // one of the two threads of the object writes its row of the A operand
            unsigned char* dst = objA + (size_t)mt * OBJA_TILE + (row >> 3) * 256 + (row & 7) * 16;
#pragma unroll
            for (int s = 0; s < AKSTEPS; ++s)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    *reinterpret_cast<float4*>(dst + s * 4096 + c * 128) =
                        make_float4(arow[s * 8 + c * 4], arow[s * 8 + c * 4 + 1], arow[s * 8 + c * 4 + 2], arow[s * 8 + c * 4 + 3]);
        }
        thr = (PASS == 2) ? P.thr2[oo] : 0.f;
        if (PASS == 1) {
            M2 = LIN ? pack2(0.f, 0.f) : pack2(-FLT_MAX, -FLT_MAX); S2 = pack2(0.f, 0.f); Rf = FLT_MAX;
            if (FUSE) {
                const float m0 = P.fz_M0[oo];
                fuse_ok = m0 > -1e30f && m0 < 1e30f && slot < P.No && o >= 0;
                if (fuse_ok) { Rf = -floorf(m0); Yc = exp2f(m0 + Rf); if (FUSE == 1) fcut = P.fz_thr * Yc; }
                fz_seg = (int)((int64_t)(blockIdx.y * TC_SPLIT + half) * P.No_pad + oo);
            }
        } else {
            const float m = P.M2[oo];
            M2 = pack2(m, m);
            acc2 = pack2(0.f, 0.f);
            if (LIN) { Rf = -m; thr = exp2f(thr - m); }      // weight cut in the linear domain: 2^(l - M) > 2^(thr - M)
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core (async proxy) reads
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == TC_CW) {
        // ===== TMA producer ======================================================================================
        if (lane == 0) {
            int lt = 0;                          // tiles actually staged (pass 2 skips the ones without a live sub-batch)
            for (int it = 0; it < nt; ++it) {
                if (PASS == 2 && P.tmask && P.tmask[(size_t)(t0 + it) * gridDim.x + blockIdx.x] == 0u) continue;
                const int st = lt % TC_NSTAGE, n = lt / TC_NSTAGE;
                ++lt;
                if (n > 0) mbar_wait_hint(&tile_empty[st], (uint32_t)((n - 1) & 1));
                mbar_expect_tx(&tile_full[st], TILE_BYTES);
                bulk_g2s(stage + (size_t)st * TILE_BYTES, tiles + (size_t)(t0 + it) * TILE_BYTES, TILE_BYTES, &tile_full[st]);
            }
        }
    } else if (warp == TC_CW + 1) {
        // ===== MMA issuer: the whole warp walks the loop (uniform control flow keeps the descriptors in uniform
        // registers), one elected lane issues ==================================================================
        // instruction descriptor: D fp32, A / B tf32, both K-major, N = TC_NC, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_NC >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t d0 = tc_desc(smem_u32(objA), lbo, sbo);
        const uint32_t a_lo = (uint32_t)d0, desc_hi = (uint32_t)(d0 >> 32);
        uint32_t gchm[TC_MT] = {0, 0};      // chunks issued per M-tile (pass 2 skips the chunks without a live sub-batch)
        int lt = 0;
        for (int it = 0; it < nt; ++it) {
            uint32_t tm = 0xffffffffu;      // live bits of the tile, OR over the objects of each M-tile (16 bits each)
            if (PASS == 2 && P.tmask) {
                tm = P.tmask[(size_t)(t0 + it) * gridDim.x + blockIdx.x];
                if (tm == 0u) continue;
            }
            const int st = lt % TC_NSTAGE, n = lt / TC_NSTAGE;
            ++lt;
            mbar_wait_hint(&tile_full[st], (uint32_t)(n & 1));
            tc_fence_after();
            const int64_t first = (t0 + it) * TC_TM;
            const int cnt = (int)((P.nm - first) < TC_TM ? (P.nm - first) : TC_TM);
            const int nch = (cnt + TC_NC - 1) / TC_NC;
            const uint32_t b_lo = (uint32_t)tc_desc(smem_u32(stage + (size_t)st * TILE_BYTES), lbo, sbo);
            for (int ch = 0; ch < nch; ++ch) {
                const uint32_t b0 = b_lo + ch * ((TC_NC / 8) * 256 >> 4);
#pragma unroll
                for (int m = 0; m < TC_MT; ++m) {
                    if (((tm >> (16 * m + 2 * ch)) & 3u) == 0u) continue;      // warp-uniform
                    const uint32_t buf = gchm[m] & 1, use = gchm[m] >> 1;
                    ++gchm[m];
                    mbar_wait_hint(&acc_empty[m * 2 + buf], (use & 1) ^ 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t dcol = tmem + buf * TC_CHUNK_COLS + m * (3 * TC_NC);
                        const uint32_t a0 = a_lo + ((m * OBJA_TILE) >> 4);
                        // inter = (-x) . m, shape = w . m^2, G = g . m: KS accumulating K = 8 instructions each; the model
                        // operand of G is that of inter
#pragma unroll
                        for (int k = 0; k < KS; ++k) {
                            const uint32_t ak = a0 + (k * 4096 >> 4), bk = b0 + (k * 8192 >> 4);
                            if (k == 0) {
                                tc_mma<false>(dcol, ak, bk, desc_hi, idesc);
                                tc_mma<false>(dcol + TC_NC, ak + (KS * 4096 >> 4), bk + (KS * 8192 >> 4), desc_hi, idesc);
                                tc_mma<false>(dcol + 2 * TC_NC, ak + (2 * KS * 4096 >> 4), bk, desc_hi, idesc);
                            } else {
                                tc_mma<true>(dcol, ak, bk, desc_hi, idesc);
                                tc_mma<true>(dcol + TC_NC, ak + (KS * 4096 >> 4), bk + (KS * 8192 >> 4), desc_hi, idesc);
                                tc_mma<true>(dcol + 2 * TC_NC, ak + (2 * KS * 4096 >> 4), bk, desc_hi, idesc);
                            }
                        }
                        tc_commit(&acc_full[m * 2 + buf]);
                    }
                    __syncwarp();
                }
            }
            if (elect_one()) tc_commit(&tile_empty[st]);     // the tensor core is done reading this stage
            __syncwarp();
        }
    } else {
        // ===== consumers =========================================================================================
        const f2 kMinusOne = pack2(-1.f, -1.f);
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + mt * (3 * TC_NC);
        int cur_bin = -1;
        uint32_t gch = 0;
        uint32_t lbits = 0;                    // pass 1: live sub-batches of the current tile; pass 2: those of the warp
        unsigned long long npairs = 0;
        auto flush = [&](float v, int bin) {
            if (v != 0.f && oidx >= 0 && (FUSE != 1 || fuse_ok)) atomicAdd(P.hist + (int64_t)oidx * P.hist_stride + bin, v);
        };
        const ulonglong2* pairs = nullptr;
        const float4* tails = nullptr;
        const int4* subs = nullptr;
        float sub_inv = 0.f;                   // pass 2, fast path: the common 1/norm of the current sub-batch
        int first_i = 0, npair_full = 0;
        bool odd = false;
        // four model pairs (eight TMEM columns) against the object of the thread
        auto process = [&](auto slow_tag, const int p0, float (&Bv)[8], float (&Cv)[8], float (&Gv)[8]) {
            constexpr bool SLOW = decltype(slow_tag)::value;
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                const int p = p0 + jp;
                bool tail = false;
                if (SLOW) {
                    tail = (p == npair_full) && odd;
                    if (p >= npair_full && !tail) continue;      // warp-uniform
                }
                f2 m2[2 * PQ];        // the bands of the two models, then the prior pair
#pragma unroll
                for (int i = 0; i < PQ; ++i) { const ulonglong2 q = pairs[p * PQ + i]; m2[2 * i] = q.x; m2[2 * i + 1] = q.y; }
                const f2 prior2 = m2[PRI];
                const int idx0 = first_i + 2 * p;
                const f2 B2 = pack2(Bv[2 * jp], Bv[2 * jp + 1]);
                const f2 C2 = pack2(Cv[2 * jp], Cv[2 * jp + 1]);
                const f2 G2 = pack2(Gv[2 * jp], Gv[2 * jp + 1]);
                f2 l;
                if (SLOW && tail) l = tc_pair_l<NF, DP, PRIOR, true, MLO>(B2, C2, G2, m2, prior2, d2, w2, K2, A2);
                else l = tc_pair_l<NF, DP, PRIOR, false, MLO>(B2, C2, G2, m2, prior2, d2, w2, K2, A2);
                const f2 delta = fma2(M2, kMinusOne, l);
                if (PASS == 1) {
                    const float e0 = fast_ex2(-fabsf(lo2(delta))), e1 = fast_ex2(-fabsf(hi2(delta)));
                    const bool g0 = lo2(delta) > 0.f, g1 = hi2(delta) > 0.f;
                    S2 = fma2(S2, pack2(g0 ? e0 : 1.f, g1 ? e1 : 1.f), pack2(g0 ? 1.f : e0, g1 ? 1.f : e1));
                    M2 = pack2(g0 ? lo2(l) : lo2(M2), g1 ? hi2(l) : hi2(M2));
                    best0 = g0 ? idx0 : best0;
                    best1 = g1 ? idx0 + 1 : best1;
                } else {
                    float u0 = fast_ex2(lo2(delta)), u1 = fast_ex2(hi2(delta));
                    const bool s0 = lo2(l) > thr, s1 = hi2(l) > thr;
                    if (P.ex_list) {   // log2 domain: a band of ex_tol / ln 2 around the cut is recorded for float64
                        const float band = P.ex_tol * 1.4427f;
                        const bool n0 = fabsf(lo2(l) - thr) < band, n1 = fabsf(hi2(l) - thr) < band && !(SLOW && tail);
                        if (__any_sync(0xffffffffu, n0 || n1)) {
                            if (n0 && oidx >= 0) record_cut(P, oidx, idx0, u0, s0);
                            if (n1 && oidx >= 0) record_cut(P, oidx, idx0 + 1, u1, s1);
                        }
                    }
                    u0 = s0 ? u0 : 0.f;
                    u1 = s1 ? u1 : 0.f;
                    if (!SLOW) {
                        acc2 = fma2(pack2(u0, u1), pack2(sub_inv, sub_inv), acc2);
                    } else {
                        const float4 tl = tails[p];
                        const int bin0 = __float_as_int(tl.z), bin1 = tail ? bin0 : __float_as_int(tl.w);
                        if (bin0 != cur_bin) {             // warp-uniform
                            if (cur_bin >= 0) { flush(lo2(acc2) + hi2(acc2), cur_bin); acc2 = pack2(0.f, 0.f); }
                            cur_bin = bin0;
                        }
                        if (bin1 == bin0) {
                            acc2 = fma2(pack2(u0, u1), pack2(tl.x, tl.y), acc2);
                        } else {                           // the pair straddles a bin boundary
                            flush(fmaf(u0, tl.x, lo2(acc2) + hi2(acc2)), bin0);
                            acc2 = pack2(0.f, u1 * tl.y);
                            cur_bin = bin1;
                        }
                    }
                }
            }
        };
        // ---- linear-domain form (LIN): with (dof/2 - 1) = 1 the likelihood is 2^l = x 2^(-x), x = chi2 log2(e)/2, so the
        // weight relative to a reference exponent R is y = x 2^(R - x + prior): one ex2 and no lg2 per pair, a plain sum
        // and a plain max.  Pass 1 keeps R per object (integer valued, so that a change of R rescales the sums by an
        // exact power of two) and lowers it whenever the largest weight passes 2^8; a jump beyond 2^24 (or an
        // overflow) has the weights of the sub-batch formed again in a frame at its minimum.  Pass 2 uses R = -M.
        // Weights of one sub-batch (four model pairs, linear domain) -> KDE histogram: those above `cut` are summed per
        // bin in a register and flushed with one RED per (thread, bin run).  A weight with |u - cmid| <= chalf is
        // recorded for the float64 re-decision (pass 2: the band of the fp32 error around the final cut, global list;
        // fused pass: from just below the running cut up to fz_gfac above it, per-thread segment).  chalf < 0: off.
        auto lin_accumulate = [&](auto slow_tag, const int p0, const f2 (&us)[4], const float cut, const float cmid,
                                  const float chalf) {
            constexpr bool SLOW = decltype(slow_tag)::value;
            float uv[8];                    // the weights that pass the cut (0 otherwise)
            float nearm = FLT_MAX;          // smallest |weight - cmid| of the sub-batch
            const f2 nmid2 = pack2(-cmid, -cmid);
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                const f2 dd = add2(us[jp], nmid2);
                nearm = fminf(nearm, fminf(fabsf(lo2(dd)), fabsf(hi2(dd))));
                uv[2 * jp] = (lo2(us[jp]) > cut) ? lo2(us[jp]) : 0.f;
                uv[2 * jp + 1] = (hi2(us[jp]) > cut) ? hi2(us[jp]) : 0.f;
            }
            if (__any_sync(0xffffffffu, nearm <= chalf)) {
                if (FUSE == 1) {
                    // the whole sub-batch (first model, cut of the moment, eight weights: 48 bytes, three vector stores)
                    // goes to the thread's segment; k_fuse_fix works out which of the weights lie in the band
                    if (nearm <= chalf && fuse_ok) {
                        if (rcnt < P.fz_cap) {
                            uint4* dst = reinterpret_cast<uint4*>(P.fz_rec) + ((size_t)fz_seg * P.fz_cap + rcnt) * 3;
                            dst[0] = make_uint4((unsigned)(first_i + 2 * p0), __float_as_uint(cut), __float_as_uint(lo2(us[0])),
                                                __float_as_uint(hi2(us[0])));
                            dst[1] = make_uint4(__float_as_uint(lo2(us[1])), __float_as_uint(hi2(us[1])), __float_as_uint(lo2(us[2])),
                                                __float_as_uint(hi2(us[2])));
                            dst[2] = make_uint4(__float_as_uint(lo2(us[3])), __float_as_uint(hi2(us[3])), 0u, 0u);
                        }
                        ++rcnt;
                    }
                } else if (nearm <= chalf && oidx >= 0) {
                    const int cnt_i = SLOW ? 2 * npair_full + (odd ? 1 : 0) : (1 << 30);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float uq = (q & 1) ? hi2(us[q >> 1]) : lo2(us[q >> 1]);
                        const int jm = 2 * (p0 + (q >> 1)) + (q & 1);
                        if (fabsf(uq - cmid) <= chalf && jm < cnt_i) record_cut(P, oidx, first_i + jm, uq, uq > cut);
                    }
                }
            }
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                const int p = p0 + jp;
                const float u0 = uv[2 * jp], u1 = uv[2 * jp + 1];
                if (!SLOW) {
                    acc2 = fma2(pack2(u0, u1), pack2(sub_inv, sub_inv), acc2);
                } else {
                    const float4 tl = tails[p];
                    if (p > npair_full || (p == npair_full && !odd)) continue;      // warp-uniform: nothing but padding
                    const bool tail = (p == npair_full);
                    const int bin0 = __float_as_int(tl.z), bin1 = tail ? bin0 : __float_as_int(tl.w);
                    if (bin0 != cur_bin) {             // warp-uniform
                        if (cur_bin >= 0) { flush(lo2(acc2) + hi2(acc2), cur_bin); acc2 = pack2(0.f, 0.f); }
                        cur_bin = bin0;
                    }
                    if (bin1 == bin0) {
                        acc2 = fma2(pack2(u0, u1), pack2(tl.x, tl.y), acc2);
                    } else {                           // the pair straddles a bin boundary
                        flush(fmaf(u0, tl.x, lo2(acc2) + hi2(acc2)), bin0);
                        acc2 = pack2(0.f, u1 * tl.y);
                        cur_bin = bin1;
                    }
                }
            }
        };
        uint32_t live_bit = 0;                 // bit of the sub-batch being processed
        auto process_lin = [&](auto slow_tag, const int p0, float (&Bv)[8], float (&Cv)[8], float (&Gv)[8]) {
            constexpr bool SLOW = decltype(slow_tag)::value;
            f2 xs[4], pr[4];
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                const int p = p0 + jp;
                f2 m2[2 * PQ];
#pragma unroll
                for (int i = 0; i < PQ; ++i) { const ulonglong2 q = pairs[p * PQ + i]; m2[2 * i] = q.x; m2[2 * i + 1] = q.y; }
                pr[jp] = m2[PRI];
                const f2 B2 = pack2(Bv[2 * jp], Bv[2 * jp + 1]);
                const f2 C2 = pack2(Cv[2 * jp], Cv[2 * jp + 1]);
                const f2 G2 = pack2(Gv[2 * jp], Gv[2 * jp + 1]);
                f2 x = tc_pair_c<NF, MLO>(B2, C2, G2, m2, d2, w2, K2);
                if (SLOW) {      // padding models (beyond the last one) get x = FLT_MAX: weight 0, never the minimum
                    const int cnt_i = 2 * npair_full + (odd ? 1 : 0);
                    x = pack2(2 * p < cnt_i ? lo2(x) : FLT_MAX, 2 * p + 1 < cnt_i ? hi2(x) : FLT_MAX);
                }
                xs[jp] = x;
            }
            if (PASS == 1) {
                // running sum S2 in the frame of R; M2 = (max weight Ym, candidate threshold Yt = (1 - 2^-18) Ym).  The
                // maximum itself is kept in the log domain (Mfl, best0), evaluated only for the candidates y > Yt, so that
                // the arg-max does not depend on the frame.  The weights of the sub-batch are formed first and applied to
                // the running state only once they are known to be representable.
                const int ig0 = first_i + 2 * p0;
                f2 ys[4], Ssub = 0;
                float ysub = 0.f;
                auto weights = [&]() {
                    const f2 R2 = pack2(Rf, Rf);
                    ysub = 0.f;
#pragma unroll
                    for (int jp = 0; jp < 4; ++jp) {
                        const f2 arg = fma2(xs[jp], kMinusOne, PRIOR ? add2(pr[jp], R2) : R2);
                        ys[jp] = mul2(xs[jp], pack2(fast_ex2(lo2(arg)), fast_ex2(hi2(arg))));
                        Ssub = jp ? add2(Ssub, ys[jp]) : ys[jp];
                        ysub = fmaxf(ysub, fmaxf(lo2(ys[jp]), hi2(ys[jp])));
                    }
                };
                weights();
                // a weight beyond 2^24 (R far above this sub-batch's chi2) or an overflow (inf / NaN in the sum): move the
                // frame of this object to the sub-batch's minimum (an exact power-of-two rescale of the running state) and
                // form the weights again
                const bool ovf = !(ysub < 16777216.f) || !(lo2(Ssub) + hi2(Ssub) < 1.2676506e30f);
                if (__any_sync(0xffffffffu, ovf)) {
                    if (ovf) {
                        float mn = FLT_MAX;
#pragma unroll
                        for (int jp = 0; jp < 4; ++jp) {
                            const f2 v = PRIOR ? fma2(pr[jp], kMinusOne, xs[jp]) : xs[jp];
                            mn = fminf(mn, fminf(lo2(v), hi2(v)));
                        }
                        // y = x 2^(R - x): R = floor(x_min) - log2(x_min) puts the largest weight near 1
                        const float ex = (mn >= 2.f && mn < 1e38f) ? (float)(((__float_as_int(mn) >> 23) & 0xff) - 127) : 0.f;
                        const float Rn = fminf(floorf(mn) - ex, Rf - 1.f);
                        const float f = (Rf < 1e38f) ? pow2i(Rn - Rf) : 0.f;
                        S2 = mul2(S2, pack2(f, f));
                        M2 = mul2(M2, pack2(f, f));
                        Sd *= (double)f;
                        Rf = Rn;
                        if (FUSE) { fuse_ok = false; fcut = FLT_MAX; Yc *= f; }      // the histogram was being filled in the old frame
                    }
                    weights();
                }
                S2 = add2(S2, Ssub);
                // may pass the final cut: pass 2 looks (FUSE: the seeded maximum makes the bits nearly the final selection)
                if (P.live && ysub > P.live_thr * (FUSE ? fmaxf(lo2(M2), Yc) : lo2(M2))) lbits |= live_bit;
                const float Yt = hi2(M2);
                if (__any_sync(0xffffffffu, ysub > Yt)) {
                    float Ym = lo2(M2);
#pragma unroll
                    for (int jp = 0; jp < 4; ++jp) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const float y = e ? hi2(ys[jp]) : lo2(ys[jp]);
                            if (y > Yt) {
                                const float x = e ? hi2(xs[jp]) : lo2(xs[jp]);
                                float l = fast_lg2(x) - x;
                                if (PRIOR) l += e ? hi2(pr[jp]) : lo2(pr[jp]);
                                if (l > Mfl) { Mfl = l; best0 = ig0 + 2 * jp + e; }
                                Ym = fmaxf(Ym, y);
                            }
                        }
                    }
                    // keep the largest weight below 2^8 (the fp32 rounding of R - x grows with |R - x|): an exact
                    // power-of-two change of frame, no recomputation
                    if (Ym >= 256.f && Ym < 1e30f) {
                        const int k = ((__float_as_int(Ym) >> 23) & 0xff) - 127;
                        const float f = __int_as_float((127 - k) << 23);
                        S2 = mul2(S2, pack2(f, f));
                        Sd *= (double)f;
                        Ym *= f;
                        Rf -= (float)k;
                        if (FUSE) { fuse_ok = false; fcut = FLT_MAX; Yc *= f; }
                    }
                    M2 = pack2(Ym, Ym * 0.99999618530273438f);
                    if (FUSE && fuse_ok && Ym > Yc) { Yc = Ym; if (FUSE == 1) fcut = P.fz_thr * Ym; }
                }
                if (FUSE == 1) {
                    // every weight above the running cut goes into the histogram now; the band from just below the cut
                    // to fz_gfac above it is recorded for k_fuse_fix (fz_mid / fz_half: centre / half-width of the band
                    // in units of the cut)
                    if (__any_sync(0xffffffffu, ysub > fcut * P.fz_lofac))
                        lin_accumulate(slow_tag, p0, ys, fcut, fcut * P.fz_mid, fcut * P.fz_half);
                }
            } else {
                const f2 R2 = pack2(Rf, Rf);
                f2 us[4];
#pragma unroll
                for (int jp = 0; jp < 4; ++jp) {
                    const f2 arg = fma2(xs[jp], kMinusOne, PRIOR ? add2(pr[jp], R2) : R2);
                    us[jp] = mul2(xs[jp], pack2(fast_ex2(lo2(arg)), fast_ex2(hi2(arg))));
                }
                // weights within the fp32 error of the cut: recorded, re-decided in float64 by k_exact_cut_fix
                lin_accumulate(slow_tag, p0, us, thr, thr, P.ex_list ? thr * P.ex_tol : -1.f);
            }
        };
        // pass 2: the live bits of the next tile are fetched while the current one is processed
        uint32_t live_next = 0;
        if (PASS == 2 && P.live && oidx >= 0 && nt > 0)
            live_next = P.live[((size_t)t0 * TC_SPLIT + half) * (size_t)P.No_pad + oidx];
        int lt = 0;                            // tiles actually staged (see the TMA producer)
        for (int it = 0; it < nt; ++it) {
            const uint32_t live_mine = live_next;
            if (PASS == 2 && P.live && oidx >= 0 && it + 1 < nt)
                live_next = P.live[((size_t)(t0 + it + 1) * TC_SPLIT + half) * (size_t)P.No_pad + oidx];
            uint32_t tmm = 0xffffu;            // live bits of the tile for this M-tile (all of its objects, both halves)
            if (PASS == 2 && P.tmask) {
                const uint32_t tm = P.tmask[(size_t)(t0 + it) * gridDim.x + blockIdx.x];
                if (tm == 0u) continue;        // neither M-tile needs the tile: it was not staged
                tmm = (tm >> (16 * mt)) & 0xffffu;
            }
            const int st = lt % TC_NSTAGE, n = lt / TC_NSTAGE;
            ++lt;
            mbar_wait_hint(&tile_full[st], (uint32_t)(n & 1));
            const unsigned char* tile = stage + (size_t)st * TILE_BYTES;
            pairs = reinterpret_cast<const ulonglong2*>(tile + OPSEC);
            tails = reinterpret_cast<const float4*>(tile + OPSEC + PAIRSEC);
            subs = reinterpret_cast<const int4*>(tile + OPSEC + PAIRSEC + TC_TAILSEC);
            const int64_t first = (t0 + it) * TC_TM;
            const int cnt = (int)((P.nm - first) < TC_TM ? (P.nm - first) : TC_TM);
            const int nch = (cnt + TC_NC - 1) / TC_NC;
            first_i = (int)first;
            npair_full = cnt >> 1;
            odd = (cnt & 1) != 0;
            const size_t live_at = ((size_t)(t0 + it) * TC_SPLIT + half) * (size_t)P.No_pad;
            if (PASS == 2) {
                // sub-batches in which some object of this warp may have a model above the cut (bits of pass 1)
                lbits = 0xffffffffu;
                if (P.live) lbits = __reduce_or_sync(0xffffffffu, (oidx >= 0) ? live_mine : 0u);
            } else {
                lbits = 0;
            }
            for (int ch = 0; ch < nch; ++ch) {
                if (((tmm >> (2 * ch)) & 3u) == 0u) continue;      // no object of the M-tile needs the chunk: no MMA was issued
                const uint32_t buf = gch & 1, use = gch >> 1;
                ++gch;
                mbar_wait_hint(&acc_full[mt * 2 + buf], use & 1);
                tc_fence_after();
                const uint32_t cbase = lane_addr + buf * TC_CHUNK_COLS;
                if (PASS == 2 && ((lbits >> (2 * ch)) & 3u) == 0u) {      // warp-uniform: nothing of this chunk survives
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[mt * 2 + buf]);
                    continue;
                }
#pragma unroll        // (rolling this loop in the fused pass to halve its code measured 6 % slower)
                for (int k = 0; k < TC_NSUB / TC_SPLIT; ++k) {
                    const int sub = half + TC_SPLIT * k;          // this warpgroup's sub-batches of the chunk
                    live_bit = 1u << (2 * ch + k);
                    const bool dead = (PASS == 2) && (lbits & live_bit) == 0u;     // warp-uniform
                    float Bv[8], Cv[8], Gv[8];
                    if (!dead) {
                        tmem_ld8(cbase + sub * 8, Bv);
                        tmem_ld8(cbase + TC_NC + sub * 8, Cv);
                        tmem_ld8(cbase + 2 * TC_NC + sub * 8, Gv);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" : TC_TIE8(Bv), TC_TIE8(Cv), TC_TIE8(Gv)::"memory");
                    }
                    if (k == TC_NSUB / TC_SPLIT - 1) {   // everything this warp needs of the chunk is in registers
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[mt * 2 + buf]);
                    }
                    if (dead) continue;
                    if (PASS == 2) npairs += 8;
                    const int p0 = ch * (TC_NC / 2) + sub * 4;
                    // fast path: four complete model pairs and (pass 2) one KDE bin for all eight models -> one
                    // branch-free block in which the chains of the four pair evaluations interleave
                    bool fast = p0 + 4 <= npair_full;
                    if (PASS == 2 || FUSE == 1) {
                        const int4 si = subs[ch * TC_NSUB + sub];
                        fast = fast && (si.y != 0);
                        sub_inv = __int_as_float(si.z);
                        if (fast && si.x != cur_bin) {         // warp-uniform
                            if (cur_bin >= 0) { flush(lo2(acc2) + hi2(acc2), cur_bin); acc2 = pack2(0.f, 0.f); }
                            cur_bin = si.x;
                        }
                    }
                    if (LIN) {
                        if (fast) process_lin(std::false_type{}, p0, Bv, Cv, Gv);
                        else process_lin(std::true_type{}, p0, Bv, Cv, Gv);
                    } else {
                        if (fast) process(std::false_type{}, p0, Bv, Cv, Gv);
                        else process(std::true_type{}, p0, Bv, Cv, Gv);
                    }
                }
            }
            if (PASS == 1 && P.live) {
                const int64_t slot = tile_base + (int64_t)mt * 128 + row;
                const int64_t wo = P.objlist ? (int64_t)oidx : slot;      // outputs are indexed by object
                if (P.objlist ? oidx >= 0 : slot < P.No_pad) P.live[live_at + wo] = (unsigned short)(LIN ? lbits : 0xffffu);
            }
            if (PASS == 1 && LIN) {
                Sd += (double)(lo2(S2) + hi2(S2));
                S2 = pack2(0.f, 0.f);
            } else if (PASS == 1) {
                // fp32 sums only within a tile; tiles (and the two model lanes) are combined in float64
                const float m0 = lo2(M2), m1 = hi2(M2);
                const float mn = fmaxf(m0, m1);
                Sd = Sd * (double)fast_ex2(Mfl - mn) + (double)(lo2(S2) * fast_ex2(m0 - mn)) + (double)(hi2(S2) * fast_ex2(m1 - mn));
                Mfl = mn;
                S2 = pack2(0.f, 0.f);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&tile_empty[st]);     // done with the pair / tail sections of this stage
        }
        if (PASS == 1) {
            const int64_t slot = tile_base + (int64_t)mt * 128 + row;
            const int64_t wo = P.objlist ? (int64_t)oidx : slot;          // outputs are indexed by object
            if (P.objlist ? oidx >= 0 : slot < P.No_pad) {
                const float m0 = lo2(M2), m1 = hi2(M2);
                // the two threads of an object report like two model splits
                const size_t q = ((size_t)blockIdx.y * TC_SPLIT + half) * P.No_pad + wo;
                if (LIN) {
                    // y = 2^(l + R): sum 2^(l - max l) = sum y / 2^(max l + R)
                    const bool any = Mfl > -FLT_MAX;
                    P.pM[q] = (double)Mfl;
                    P.pS[q] = any ? Sd * exp2(-((double)Mfl + (double)Rf)) : 0.0;
                    P.pbest[q] = best0;
                    if (FUSE == 1) {
                        if (cur_bin >= 0) flush(lo2(acc2) + hi2(acc2), cur_bin);
                        P.fz_cnt[q] = fuse_ok ? rcnt : -1;
                    } else if (FUSE == 2) {
                        P.fz_cnt[q] = -1;
                    }
                } else {
                    P.pM[q] = (double)Mfl;
                    P.pS[q] = Sd;
                    P.pbest[q] = (m0 > m1) ? best0 : ((m1 > m0) ? best1 : min(best0, best1));
                }
            }
        } else {
            if (cur_bin >= 0) flush(lo2(acc2) + hi2(acc2), cur_bin);
            if (P.pairs_done) {
                const unsigned long long mine = (oidx >= 0) ? npairs : 0ull;
                unsigned long long tot = mine;
                for (int sft = 16; sft > 0; sft >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, sft);
                if (lane == 0 && tot) atomicAdd(P.pairs_done, tot);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_CW + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS));
}

// ---- tile builder ------------------------------------------------------------------------------------------
struct TcRecParams {
    const double* m;
    const double* lnprior;
    const int32_t* perm;
    const int32_t* bins;
    const float* invnorm;
    int64_t nm;
    int Nf;
    int mlo;
    unsigned char* tiles;
    int stride;              // tile position p holds the model at sorted position p * stride (coarse tile set)
};

__global__ void k_build_tiles_tc(TcRecParams P) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.nm) return;
    const int64_t ps = p * P.stride;          // sorted position
    const int64_t j = P.perm[ps];
    const int r = (int)(p % TC_TM);
    const bool mlo = P.mlo != 0;
    const int nf = P.Nf, slot = tc_slot(nf), ks = tc_ks(nf), pq = tc_pairq(nf, mlo);
    const int opsec = tc_opsec(nf), pairsec = tc_pairsec(nf, mlo);
    unsigned char* T = P.tiles + (size_t)(p / TC_TM) * tc_tile_bytes(nf, mlo);
    float rowv[8 * 6];
    for (int i = 0; i < 8 * 6; ++i) rowv[i] = 0.f;
    float* pr = reinterpret_cast<float*>(T + opsec) + (r >> 1) * (4 * pq);
    for (int b = 0; b < nf; ++b) {
        const double v = P.m[j * nf + b];
        const float mf = (float)v, q = (float)(v * v);
        const float mh = tf32_rn(mf), ml = tf32_rn(mf - mh);
        const float qh = tf32_rn(q), ql = tf32_rn(q - qh);
        rowv[b] = mh; rowv[slot + b] = ml; rowv[2 * slot + b] = mh;
        rowv[8 * ks + b] = qh; rowv[8 * ks + slot + b] = ql; rowv[8 * ks + 2 * slot + b] = qh;
        pr[2 * b + (r & 1)] = mf;
        if (mlo) pr[2 * (nf + b) + (r & 1)] = (float)(v - (double)mf);
    }
    pr[2 * (mlo ? 2 * nf : nf) + (r & 1)] = P.lnprior ? (float)(P.lnprior[j] * 1.4426950408889634) : 0.f;
    float* tl = reinterpret_cast<float*>(T + opsec + pairsec) + (r >> 1) * 4;
    const bool kde = P.bins != nullptr && P.stride == 1;      // the coarse set serves pass 1 only
    tl[r & 1] = kde ? P.invnorm[p] : 0.f;
    tl[2 + (r & 1)] = __int_as_float(kde ? P.bins[p] : -1);
    if ((r & 7) == 0) {
        int uniform = (kde && p + 8 <= P.nm) ? 1 : 0;
        const int b0 = kde ? P.bins[p] : -1;
        for (int i = 1; i < 8 && uniform; ++i) uniform = (P.bins[p + i] == b0) ? 1 : 0;
        reinterpret_cast<int4*>(T + opsec + pairsec + TC_TAILSEC)[r >> 3] =
            make_int4(b0, uniform, __float_as_int(kde ? P.invnorm[p] : 0.f), 0);
    }
    unsigned char* dst = T + (r >> 3) * 256 + (r & 7) * 16;
    for (int s = 0; s < 2 * ks; ++s)
        for (int c = 0; c < 2; ++c)
            *reinterpret_cast<float4*>(dst + s * (TC_TM * 32) + c * 128) =
                make_float4(rowv[s * 8 + c * 4], rowv[s * 8 + c * 4 + 1], rowv[s * 8 + c * 4 + 2], rowv[s * 8 + c * 4 + 3]);
}
