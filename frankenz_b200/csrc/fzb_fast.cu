// fp32 register-tiled path for the reduce-only shapes (fit_predict with save_fits=False).
//
// Layout: one thread owns R objects (photometry, weights and the running reduction state live in
// registers); all threads of a CTA walk the same model tile, which a single elected thread stages
// into shared memory with TMA bulk copies (cp.async.bulk + mbarrier, double buffered), so every
// model value is a shared-memory broadcast.  The (N_obj x N_model) matrix never exists:
//
//   pass 1  k_sweep<PASS=1>  per pair: chi2 / optimal scale -> ln-likelihood (log2 units) ->
//                            online (max, sum 2^(l-max), argmax) per object
//   merge   k_merge          combine model splits, re-evaluate the best pair exactly in float64
//                            (lmap, chi2, scale), levid = lmap + ln(sum), decide per object
//                            whether the fp32 error bound allows the fp32 pass 2
//   pass 2  k_sweep<PASS=2>  recompute l, weight = 2^(l - max) if above the wt_thresh cut, sum the
//                            weights of each run of models that share a KDE bin in a register and
//                            flush one RED per (object, bin) into a global histogram
//   finish  k_finish         histogram (*) tabulated Gaussian kernel in float64, normalise, write PDF
//
// Models are sorted by KDE bin (dictionary width, grid position) when the labels are set, which is
// what makes the "one flush per bin" accumulation possible.  Objects whose fp32 result cannot be
// trusted to the 1e-5 parity bound (large best-fit chi2, extreme S/N, degenerate rows, non-finite
// sums) are routed to the float64 kernels of fzb_generic.cu.
//
// Precision: inputs are split hi/lo (float64 = hi + lo).  chi2 is evaluated on the hi parts and
// corrected to first order for the lo parts, d(chi2) = 2 sum_b w_b r_b (dlo_b - s mlo_b), which is
// exact to second order in 2^-24 (envelope theorem: the optimal scale need not be re-derived).
// This removes the loss of the low bits in the cancellation d - s*m for bright objects
// (hard part 1) at 5 extra FMAs per pair (10 when the models are not
// exactly representable in fp32).
#include <algorithm>
#include <cfloat>
#include <float.h>
#include <cstdlib>
#include <numeric>
#include <type_traits>

#include "fzb_common.cuh"
#include "fzb_pair64.cuh"

#include "fzb_sweep_common.cuh"

namespace {

using namespace fzbsweep;

// =====================================================================================================
// Packed-FP32 sweep (FFMA2 / FMUL2 / FADD2): two objects per 64-bit register pair.
// `fma.rn.f32x2` issues at half the instruction rate of FFMA (tools/peak_ffma2.cu) but does two lanes'
// worth of work, so the 30 FMA-class operations of a pair cost 15 issue slots instead of 30; the freed
// slots go to the MUFU / select / shared-memory instructions that otherwise compete with the FMAs.
// Model values are stored duplicated (m, m) in the shared-memory record so one LDS.128 yields two
// ready-made packed operands.
// =====================================================================================================
// threads per CTA / CTAs per SM of the packed kernel: R=4 -> 384 threads x 1 (<= 168 registers), R=2 -> 256 threads x 3 (<= 85)
__host__ __device__ constexpr int ft2_of(int R) { return R >= 4 ? 384 : 256; }
__host__ __device__ constexpr int minb2_of(int R) { return R >= 4 ? 1 : 3; }

__host__ __device__ constexpr int rec2_floats(int nf, int mode, bool mlo, bool mm = false) {
    // pairs (v, v): m[nf], aux[nf] (FS0: m^2, FX1: err^2), ml[nf] (MLO), k[nf] (MM: band mask 0/1);
    // tail: prior2 x2, invnorm x2, bin, (MM: mask bits)
    int n = 2 * nf * (1 + ((mode == FM_FX0) ? 0 : 1) + (mlo ? 1 : 0) + (mm ? 1 : 0)) + 5 + (mm ? 1 : 0);
    return (n + 3) / 4 * 4;
}

template <int NF, int MODE>
struct ObjPack {     // two objects (lo half, hi half)
    f2 d[NF];
    f2 w[NF];        // FS0/FX0: weights; FX1: err^2 (+inf where masked)
    f2 x[(MODE == FM_FS0) ? NF : 1];
    f2 dl[NF];       // 2 * d_lo
    f2 A;
};

template <int NF, int MODE, bool MLO, bool MM = false>
__device__ __forceinline__ f2 pack_chi2(const ObjPack<NF, MODE>& o, const f2* __restrict__ m,
                                        const f2* __restrict__ aux, const f2* __restrict__ ml,
                                        const f2* __restrict__ km = nullptr) {
    f2 ns = 0;   // minus the optimal scale (FS0)
    if (MODE == FM_FS0) {
        f2 inter = mul2(o.x[0], m[0]);
        f2 shape = mul2(o.w[0], aux[0]);
#pragma unroll
        for (int b = 1; b < NF; ++b) {
            inter = fma2(o.x[b], m[b], inter);
            shape = fma2(o.w[b], aux[b], shape);
        }
        float n0 = __fmul_rn(-lo2(inter), fast_rcp(lo2(shape)));
        float n1 = __fmul_rn(-hi2(inter), fast_rcp(hi2(shape)));
        ns = pack2(n0, n1);
    }
    f2 chi2 = 0, cm = 0;
#pragma unroll
    for (int b = 0; b < NF; ++b) {
        f2 r, w;
        if (MODE == FM_FS0) r = fma2(ns, m[b], o.d[b]);
        else r = add2(o.d[b], m[b]);            // fixed scale: the record stores -m
        if (MODE == FM_FX1) {
            f2 var = add2(o.w[b], aux[b]);
            w = pack2(fast_rcp(lo2(var)), fast_rcp(hi2(var)));
        } else {
            w = o.w[b];
        }
        if (MM) w = mul2(w, km[b]);             // model band mask (the record stores m = m^2 = 0 for masked bands)
        f2 t = mul2(r, w);
        f2 u = add2(r, o.dl[b]);                // chi2 + first-order lo correction: sum t * (r + 2 d_lo)
        if (b == 0) {
            chi2 = mul2(t, u);
            if (MLO) cm = mul2(t, ml[0]);
        } else {
            chi2 = fma2(t, u, chi2);
            if (MLO) cm = fma2(t, ml[b], cm);
        }
    }
    if (MLO) chi2 = (MODE == FM_FS0) ? fma2(ns, cm, chi2) : add2(chi2, cm);   // FX: record stores -2*m_lo
    return chi2;
}

// FUSE (pass 1): the single-pass variant, as in the tensor-core sweep (fzb_sweep_tc.cuh) but in the log domain.  The running
// maximum starts at M0, a lower bound of the object's maximum from a pre-pass over every 16th model, so the sums, the live
// bits and the selection all refer to max(M0, maximum so far): every weight above the RUNNING cut is added to a per-bin
// register sum kept in the frame of the current maximum (rescaled by the same factor as the evidence sum when the maximum
// moves; flushed to the histogram in the fixed frame of M0), and weights between just below the cut and fz_gfac above it are
// recorded (CutRecord list) for the float64 re-decision.  k_merge accepts the histogram when the final maximum is within the
// band above M0; the other objects take the pruned pass 2.
template <int NF, int MODE, bool DP, bool MLO, bool PRIOR, int R, int PASS, bool MM = false, bool FUSE = false>
__global__ void __launch_bounds__(ft2_of(R), minb2_of(R)) k_sweep2(SweepParams P) {
    static_assert(R % 2 == 0, "objects come in packed pairs");
    static_assert(!FUSE || (PASS == 1 && !MM), "the fused variant is pass 1 without model masks");
    constexpr int FT2 = ft2_of(R);
    constexpr int NP = R / 2;
    constexpr int REC = rec2_floats(NF, MODE, MLO, MM);
    constexpr int AUXOFF = 2 * NF;
    constexpr int MLOFF = 2 * NF * (1 + ((MODE == FM_FX0) ? 0 : 1));
    constexpr int KMOFF = MLOFF + (MLO ? 2 * NF : 0);      // band-mask pairs (MM)
    constexpr int TAILOFF = KMOFF + (MM ? 2 * NF : 0);     // prior2 x2, invnorm x2, bin, (mask bits)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NSTAGE * TM * REC * sizeof(float));
    // MM: the dimensionality of a pair is popc(object bits & model bits); (dof/2 - 1) and the ndim-dependent constant
    // of the ln-likelihood come from two small shared-memory tables
    float* sA = reinterpret_cast<float*>(bars + NSTAGE);
    float* sK = sA + 16;
    const int tid = threadIdx.x;
    if (MM && tid < 16) { sA[tid] = P.Atab[tid]; sK[tid] = P.Ktab[tid]; }
    int obits[R];

    ObjPack<NF, MODE> ob[NP];
    int oidx[R];
    f2 M[NP];             // running (pass 1) / final (pass 2) maximum
    f2 S[NP];             // pass 1: running sum.  pass 2: unused
    float thr[R];         // pass 2: selection cut
    int best[R];
    f2 acc[NP];
    // pass 1 sums 2^(l - max) in fp32 only within one model tile; tiles are combined in float64 so that the
    // rounding error of the evidence does not grow with the number of models
    double Sd[R];
    float Mfl[R];
    // pass-2 pruning (same idea as the tensor-core sweep, one bit per object and 256-model tile): pass 1 notes the tiles in
    // which some weight exceeds wt_thresh x the running maximum (a superset of the final selection); pass 2, whose objects are
    // sorted by live signature, skips a tile when none of the 32 x R objects of a warp has the bit.  The reference's default
    // likelihood on training rows (C1 / C2 / C5) gives narrow posteriors: most tiles are dead.
    bool lv[R];
    const float lthr = P.live_lthr;           // log2 of the cut, relative to the running maximum
    float m0f[R];                             // FUSE: the seed of the running maximum = frame of the histogram
    const int64_t tile_base = (int64_t)blockIdx.x * (FT2 * R);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        int64_t oo[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int r = 2 * p + h;
            int64_t slot = tile_base + (int64_t)r * FT2 + tid;
            int64_t o;
            if (PASS == 1) o = slot < P.No_pad ? slot : P.No_pad - 1;
            else o = slot < P.No ? P.objlist[slot] : -1;
            oidx[r] = (int)o;
            oo[h] = o < 0 ? 0 : o;
            best[r] = 0;
            Sd[r] = 0.0;
            Mfl[r] = -FLT_MAX;
            thr[r] = (PASS == 2) ? P.thr2[oo[h]] : 0.f;
            obits[r] = MM ? P.obits[oo[h]] : 0;
        }
#pragma unroll
        for (int b = 0; b < NF; ++b) {
            ob[p].d[b] = pack2(P.od[b * P.No_pad + oo[0]], P.od[b * P.No_pad + oo[1]]);
            ob[p].w[b] = pack2(P.ow[b * P.No_pad + oo[0]], P.ow[b * P.No_pad + oo[1]]);
            ob[p].dl[b] = pack2(P.odl[b * P.No_pad + oo[0]], P.odl[b * P.No_pad + oo[1]]);
            if (MODE == FM_FS0) ob[p].x[b] = pack2(P.ox[b * P.No_pad + oo[0]], P.ox[b * P.No_pad + oo[1]]);
        }
        ob[p].A = pack2(P.oA[oo[0]], P.oA[oo[1]]);
        if (PASS == 1) {
            M[p] = pack2(-FLT_MAX, -FLT_MAX); S[p] = pack2(0.f, 0.f);
            if (FUSE) {
                m0f[2 * p] = P.fz_M0[oo[0]]; m0f[2 * p + 1] = P.fz_M0[oo[1]];     // -FLT_MAX: no seed (k_merge rejects the object)
                M[p] = pack2(m0f[2 * p], m0f[2 * p + 1]);
                acc[p] = pack2(0.f, 0.f);
            }
        } else { M[p] = pack2(P.M2[oo[0]], P.M2[oo[1]]); acc[p] = pack2(0.f, 0.f); }
    }

    const int64_t ntiles_all = (P.nm + TM - 1) / TM;
    const int64_t t0 = (int64_t)blockIdx.y * P.tiles_per_split;
    int64_t t1 = t0 + P.tiles_per_split;
    if (t1 > ntiles_all) t1 = ntiles_all;
    const int nt = (int)(t1 - t0);
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int it) {
        int64_t first = (t0 + it) * TM;
        int cnt = (int)((P.nm - first) < TM ? (P.nm - first) : TM);
        uint32_t bytes = (uint32_t)cnt * REC * sizeof(float);
        uint64_t* bar = &bars[it % NSTAGE];
        mbar_expect_tx(bar, bytes);
        bulk_g2s(stage + (size_t)(it % NSTAGE) * TM * REC, P.recs + first * REC, bytes, bar);
    };
    if (tid == 0) {
        for (int it = 0; it < NSTAGE && it < nt; ++it) issue(it);
    }
    const f2 kNegHalfLog2e = pack2(-kHalfLog2e, -kHalfLog2e);
    const f2 kMinusOne = pack2(-1.f, -1.f);
    int cur_bin = -1;
    unsigned long long npairs = 0;            // pass 2: models this warp really evaluated (statistics)
    unsigned int live_next = 0;
    if (PASS == 2 && P.live && nt > 0) {
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (oidx[r] >= 0) live_next |= P.live[(size_t)t0 * (size_t)P.No_pad + oidx[r]];
    }
    for (int it = 0; it < nt; ++it) {
        const int st = it % NSTAGE;
        bool skip = false;
        if (PASS == 2 && P.live) {
            // warp-uniform: no object of this warp has a weight above the cut in this tile
            skip = __reduce_or_sync(0xffffffffu, live_next) == 0u;
            live_next = 0;
            if (it + 1 < nt) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (oidx[r] >= 0) live_next |= P.live[(size_t)(t0 + it + 1) * (size_t)P.No_pad + oidx[r]];
            }
        }
        if (PASS == 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) lv[r] = false;
        }
        mbar_wait(&bars[st], (uint32_t)((it / NSTAGE) & 1));
        const float* tile = stage + (size_t)st * TM * REC;
        const int64_t first = (t0 + it) * TM;
        const int cnt = skip ? 0 : (int)((P.nm - first) < TM ? (P.nm - first) : TM);
        if (PASS == 2 && !skip) npairs += cnt;
#pragma unroll 2
        for (int jj = 0; jj < cnt; ++jj) {
            const float* rec = tile + jj * REC;       // same record for every thread: shared-memory broadcast
            const f2* rec2 = reinterpret_cast<const f2*>(rec);
            const f2 prior2 = PRIOR ? rec2[TAILOFF / 2] : 0;
            f2 invnorm = 0;
            if (PASS == 2 || FUSE) {
                invnorm = rec2[TAILOFF / 2 + 1];
                const int bin = __float_as_int(rec[TAILOFF + 4]);
                if (bin != cur_bin) {                  // warp-uniform
                    if (cur_bin >= 0) {
#pragma unroll
                        for (int p = 0; p < NP; ++p) {
                            float a0 = lo2(acc[p]), a1 = hi2(acc[p]);
                            if (FUSE) {      // register sums are in the frame of the current maximum, the histogram in that of M0
                                if (a0 != 0.f) a0 *= fast_ex2(lo2(M[p]) - m0f[2 * p]);
                                if (a1 != 0.f) a1 *= fast_ex2(hi2(M[p]) - m0f[2 * p + 1]);
                            }
                            if (a0 != 0.f && oidx[2 * p] >= 0)
                                atomicAdd(P.hist + (int64_t)oidx[2 * p] * P.hist_stride + cur_bin, a0);
                            if (a1 != 0.f && oidx[2 * p + 1] >= 0)
                                atomicAdd(P.hist + (int64_t)oidx[2 * p + 1] * P.hist_stride + cur_bin, a1);
                            acc[p] = pack2(0.f, 0.f);
                        }
                    }
                    cur_bin = bin;
                }
            }
            f2 m[NF], aux[NF], ml[NF], km[NF];
#pragma unroll
            for (int b = 0; b < NF; ++b) {
                m[b] = rec2[b];
                aux[b] = (MODE == FM_FX0) ? 0 : rec2[AUXOFF / 2 + b];
                ml[b] = MLO ? rec2[MLOFF / 2 + b] : 0;
                km[b] = MM ? rec2[KMOFF / 2 + b] : 0;
            }
            const int mbits = MM ? __float_as_int(rec[TAILOFF + 5]) : 0;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                f2 chi2 = pack_chi2<NF, MODE, MLO, MM>(ob[p], m, aux, ml, km);
                // ln-likelihood in log2 units up to a per-object constant.  For FS0 / FX0 the factor -log2(e)/2 is
                // folded into the object weights (k_prep_objects), so `chi2` already holds c = -chi2*log2(e)/2 and
                // l = (dof/2-1) log2|c| + c (+ prior): one packed FMA.  FX1 weights are per pair, so c is formed here.
                f2 c = (MODE == FM_FX1) ? mul2(chi2, kNegHalfLog2e) : chi2;
                if (PRIOR) c = add2(c, prior2);
                f2 l = c;
                if (MM) {
                    const int n0 = __popc(obits[2 * p] & mbits), n1 = __popc(obits[2 * p + 1] & mbits);
                    f2 cc = (MODE == FM_FX1) ? mul2(chi2, kNegHalfLog2e) : chi2;
                    // a free-scale fit through a single common band is exact: chi2 = 0 in the reference (float64),
                    // which then produces inf / NaN under dim_prior; fp32 leaves a rounding residue, so force it
                    if (MODE == FM_FS0) cc = pack2(n0 <= 1 ? 0.f : lo2(cc), n1 <= 1 ? 0.f : hi2(cc));
                    c = add2(add2(cc, prior2), pack2(sK[n0], sK[n1]));
                    l = c;
                    if (DP)
                        l = fma2(pack2(sA[n0], sA[n1]), pack2(fast_lg2(fabsf(lo2(cc))), fast_lg2(fabsf(hi2(cc)))), c);
                } else if (DP) {
                    f2 cc = (MODE == FM_FX1 || PRIOR) ? ((MODE == FM_FX1) ? mul2(chi2, kNegHalfLog2e) : chi2) : c;
                    l = fma2(ob[p].A, pack2(fast_lg2(fabsf(lo2(cc))), fast_lg2(fabsf(hi2(cc)))), c);
                }
                f2 delta = fma2(M[p], kMinusOne, l);
                float d0 = lo2(delta), d1 = hi2(delta);
                if (PASS == 1) {
                    float e0 = fast_ex2(-fabsf(d0)), e1 = fast_ex2(-fabsf(d1));
                    bool g0 = d0 > 0.f, g1 = d1 > 0.f;
                    S[p] = fma2(S[p], pack2(g0 ? e0 : 1.f, g1 ? e1 : 1.f), pack2(g0 ? 1.f : e0, g1 ? 1.f : e1));
                    M[p] = pack2(g0 ? lo2(l) : lo2(M[p]), g1 ? hi2(l) : hi2(M[p]));
                    best[2 * p] = g0 ? (int)(first + jj) : best[2 * p];
                    best[2 * p + 1] = g1 ? (int)(first + jj) : best[2 * p + 1];
                    lv[2 * p] = lv[2 * p] || d0 > lthr;
                    lv[2 * p + 1] = lv[2 * p + 1] || d1 > lthr;
                    if (FUSE) {
                        // above the running cut (a new maximum always is); P.fz_thr = log2(wt_thresh) in this kernel
                        const bool s0 = d0 > P.fz_thr, s1 = d1 > P.fz_thr;
                        // weight relative to the maximum AFTER this model: 1 for a new maximum, e otherwise; the sum so far is
                        // rescaled like the evidence sum
                        const float w0 = g0 ? 1.f : e0, w1 = g1 ? 1.f : e1;
                        acc[p] = fma2(acc[p], pack2(g0 ? e0 : 1.f, g1 ? e1 : 1.f), mul2(pack2(s0 ? w0 : 0.f, s1 ? w1 : 0.f), invnorm));
                        const bool n0 = fabsf(d0 - P.fz_mid) <= P.fz_half, n1 = fabsf(d1 - P.fz_mid) <= P.fz_half;
                        if (n0 || n1) {      // rare: in the band around the running cut -> float64 re-decision (weight in the frame of M0)
                            if (n0 && oidx[2 * p] >= 0 && m0f[2 * p] > -1e30f)
                                record_cut(P, oidx[2 * p], (int)(first + jj), fast_ex2(lo2(l) - m0f[2 * p]), s0);
                            if (n1 && oidx[2 * p + 1] >= 0 && m0f[2 * p + 1] > -1e30f)
                                record_cut(P, oidx[2 * p + 1], (int)(first + jj), fast_ex2(hi2(l) - m0f[2 * p + 1]), s1);
                        }
                    }
                } else {
                    float u0 = fast_ex2(d0), u1 = fast_ex2(d1);
                    const bool s0 = lo2(l) > thr[2 * p], s1 = hi2(l) > thr[2 * p + 1];
                    if (P.ex_list) {   // weights within the fp32 error of the cut: recorded for k_exact_cut_fix
                        const float band = P.ex_tol * 1.4427f;
                        const bool n0 = fabsf(lo2(l) - thr[2 * p]) < band, n1 = fabsf(hi2(l) - thr[2 * p + 1]) < band;
                        if (n0 && oidx[2 * p] >= 0) record_cut(P, oidx[2 * p], (int)(first + jj), u0, s0);
                        if (n1 && oidx[2 * p + 1] >= 0) record_cut(P, oidx[2 * p + 1], (int)(first + jj), u1, s1);
                    }
                    u0 = s0 ? u0 : 0.f;
                    u1 = s1 ? u1 : 0.f;
                    acc[p] = fma2(pack2(u0, u1), invnorm, acc[p]);
                }
            }
        }
        if (PASS == 1) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                float m0 = lo2(M[p]), m1 = hi2(M[p]);
                Sd[2 * p] = Sd[2 * p] * (double)fast_ex2(Mfl[2 * p] - m0) + (double)lo2(S[p]);
                Sd[2 * p + 1] = Sd[2 * p + 1] * (double)fast_ex2(Mfl[2 * p + 1] - m1) + (double)hi2(S[p]);
                Mfl[2 * p] = m0;
                Mfl[2 * p + 1] = m1;
                S[p] = pack2(0.f, 0.f);
            }
            if (P.live) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int64_t slot = tile_base + (int64_t)r * FT2 + tid;
                    if (slot < P.No_pad) P.live[(size_t)(t0 + it) * (size_t)P.No_pad + slot] = lv[r] ? 1 : 0;
                }
            }
        }
        __syncthreads();
        if (tid == 0 && it + NSTAGE < nt) issue(it + NSTAGE);
    }
    if (PASS == 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int64_t slot = tile_base + (int64_t)r * FT2 + tid;
            if (slot < P.No_pad) {
                size_t q = (size_t)blockIdx.y * P.No_pad + slot;
                P.pM[q] = (double)Mfl[r];
                P.pS[q] = Sd[r];
                P.pbest[q] = best[r];
            }
        }
        if (FUSE && cur_bin >= 0) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                float a0 = lo2(acc[p]), a1 = hi2(acc[p]);
                if (a0 != 0.f && oidx[2 * p] < P.No)
                    atomicAdd(P.hist + (int64_t)oidx[2 * p] * P.hist_stride + cur_bin, a0 * fast_ex2(lo2(M[p]) - m0f[2 * p]));
                if (a1 != 0.f && oidx[2 * p + 1] < P.No)
                    atomicAdd(P.hist + (int64_t)oidx[2 * p + 1] * P.hist_stride + cur_bin, a1 * fast_ex2(hi2(M[p]) - m0f[2 * p + 1]));
            }
        }
    } else {
        if (cur_bin >= 0) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                float a0 = lo2(acc[p]), a1 = hi2(acc[p]);
                if (a0 != 0.f && oidx[2 * p] >= 0) atomicAdd(P.hist + (int64_t)oidx[2 * p] * P.hist_stride + cur_bin, a0);
                if (a1 != 0.f && oidx[2 * p + 1] >= 0)
                    atomicAdd(P.hist + (int64_t)oidx[2 * p + 1] * P.hist_stride + cur_bin, a1);
            }
        }
        if (P.pairs_done) {
            int nobj = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) nobj += oidx[r] >= 0 ? 1 : 0;
            unsigned long long tot = npairs * (unsigned long long)nobj;
            for (int sft = 16; sft > 0; sft >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, sft);
            if ((tid & 31) == 0 && tot) atomicAdd(P.pairs_done, tot);
        }
    }
}




// =====================================================================================================
// Float64 register-tiled sweep: same structure as k_sweep2 (objects in registers, model tiles staged by
// TMA bulk copies, online reductions) with chi2 evaluated in double precision, for the objects whose
// best-fit chi2 is too large for fp32 (template mismatch, very bright objects).  Only log2 / exp2 go
// through the fp32 special-function unit: their absolute error (~1e-7) does not scale with chi2.
// B200 issues DFMA at half the FP32 rate, so this path runs at a sizeable fraction of the fp32 sweep
// instead of the L2-bound generic kernel.
// =====================================================================================================
constexpr int FT64 = 256;
constexpr int R64 = 2;

__host__ __device__ constexpr int rec64_doubles(int nf, int mode) {
    int n = nf * ((mode == FM_FX0) ? 1 : 2) + 2;      // m, aux, prior2, {bin, invnorm}
    return (n + 1) / 2 * 2;
}

struct Sweep64Params {
    const double *x, *xe, *xm;      // raw objects of this chunk (No x Nf)
    int64_t No, No_pad;
    const int32_t* objlist;         // chunk-local object indices
    int64_t nlist;
    int free_scale, dim_prior;
    const double* recs;             // [nm][REC64]
    int64_t nm;
    int tiles_per_split;
    double* pM;                     // pass 1 out: [nsplit][No_pad] by object, or [nsplit][part_stride] by list position
    int64_t part_stride;            // > 0: partials indexed by position in objlist (compact, allows many splits)
    double* pS;
    int32_t* pbest;
    const double* M2;               // [No_pad]  pass 2 in
    const double* thr2;
    float* hist;
    int64_t hist_stride;
    // pass-2 pruning by whole tiles, as in k_sweep2: [model tile][No_pad] by object, one bit; null: off
    unsigned short* live;
    double live_lthr;
};

template <int NF, int MODE, bool DP, int PASS>
__global__ void __launch_bounds__(FT64, 2) k_sweep64(Sweep64Params P) {
    constexpr int REC = rec64_doubles(NF, MODE);
    constexpr int AUXOFF = NF;
    constexpr int TAILOFF = NF * ((MODE == FM_FX0) ? 1 : 2);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage = reinterpret_cast<double*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NSTAGE * TM * REC * sizeof(double));
    const int tid = threadIdx.x;
    double d[R64][NF], w[R64][NF], xw[R64][NF], A[R64], M[R64], Sd[R64], Mfl[R64], thr[R64];
    float S[R64], acc[R64];
    int oidx[R64], best[R64];
#pragma unroll
    for (int r = 0; r < R64; ++r) {
        int64_t slot = (int64_t)blockIdx.x * (FT64 * R64) + (int64_t)r * FT64 + tid;
        int o = slot < P.nlist ? P.objlist[slot] : -1;
        oidx[r] = o;
        int64_t oo = o < 0 ? 0 : o;
        double ndim = 0.0;
#pragma unroll
        for (int b = 0; b < NF; ++b) {
            double v = P.x[oo * NF + b], e = P.xe[oo * NF + b], k = P.xm[oo * NF + b];
            bool clean = isfinite(v) && isfinite(e) && (e > 0.0);
            if (!clean) { v = 0.0; e = 1.0; k = 0.0; }
            d[r][b] = v;
            if (MODE == FM_FX1) w[r][b] = (k != 0.0) ? e * e : CUDART_INF;
            else w[r][b] = k / (e * e);
            xw[r][b] = v * w[r][b];
            ndim += k;
        }
        double a = P.free_scale ? 0.5 * (ndim - 1.0) : 0.5 * ndim;
        A[r] = P.dim_prior ? a - 1.0 : 0.0;
        S[r] = 0.f; Sd[r] = 0.0; Mfl[r] = -DBL_MAX; best[r] = 0; acc[r] = 0.f;
        if (PASS == 1) { M[r] = -DBL_MAX; thr[r] = 0.0; }
        else { M[r] = P.M2[oo]; thr[r] = P.thr2[oo]; }
    }
    const int64_t ntiles_all = (P.nm + TM - 1) / TM;
    const int64_t t0 = (int64_t)blockIdx.y * P.tiles_per_split;
    int64_t t1 = t0 + P.tiles_per_split;
    if (t1 > ntiles_all) t1 = ntiles_all;
    const int nt = (int)(t1 - t0);
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int it) {
        int64_t first = (t0 + it) * TM;
        int cnt = (int)((P.nm - first) < TM ? (P.nm - first) : TM);
        uint32_t bytes = (uint32_t)cnt * REC * sizeof(double);
        uint64_t* bar = &bars[it % NSTAGE];
        mbar_expect_tx(bar, bytes);
        bulk_g2s(stage + (size_t)(it % NSTAGE) * TM * REC, P.recs + first * REC, bytes, bar);
    };
    if (tid == 0) {
        for (int it = 0; it < NSTAGE && it < nt; ++it) issue(it);
    }
    int cur_bin = -1;
    bool lv[R64];
    unsigned int live_next = 0;
    if (PASS == 2 && P.live && nt > 0) {
#pragma unroll
        for (int r = 0; r < R64; ++r)
            if (oidx[r] >= 0) live_next |= P.live[(size_t)t0 * (size_t)P.No_pad + oidx[r]];
    }
    for (int it = 0; it < nt; ++it) {
        const int st = it % NSTAGE;
        bool skip = false;
        if (PASS == 2 && P.live) {       // warp-uniform: no object of this warp has a weight above the cut in this tile
            skip = __reduce_or_sync(0xffffffffu, live_next) == 0u;
            live_next = 0;
            if (it + 1 < nt) {
#pragma unroll
                for (int r = 0; r < R64; ++r)
                    if (oidx[r] >= 0) live_next |= P.live[(size_t)(t0 + it + 1) * (size_t)P.No_pad + oidx[r]];
            }
        }
        if (PASS == 1) {
#pragma unroll
            for (int r = 0; r < R64; ++r) lv[r] = false;
        }
        mbar_wait(&bars[st], (uint32_t)((it / NSTAGE) & 1));
        const double* tile = stage + (size_t)st * TM * REC;
        const int64_t first = (t0 + it) * TM;
        const int cnt = skip ? 0 : (int)((P.nm - first) < TM ? (P.nm - first) : TM);
#pragma unroll 1
        for (int jj = 0; jj < cnt; ++jj) {
            const double* rec = tile + jj * REC;
            const double prior2 = rec[TAILOFF];
            float invnorm = 0.f;
            if (PASS == 2) {
                const int2 tl = *reinterpret_cast<const int2*>(rec + TAILOFF + 1);
                invnorm = __int_as_float(tl.y);
                if (tl.x != cur_bin) {
                    if (cur_bin >= 0) {
#pragma unroll
                        for (int r = 0; r < R64; ++r) {
                            if (acc[r] != 0.f && oidx[r] >= 0)
                                atomicAdd(P.hist + (int64_t)oidx[r] * P.hist_stride + cur_bin, acc[r]);
                            acc[r] = 0.f;
                        }
                    }
                    cur_bin = tl.x;
                }
            }
            double m[NF], aux[NF];
#pragma unroll
            for (int b = 0; b < NF; ++b) {
                m[b] = rec[b];
                aux[b] = (MODE == FM_FX0) ? 0.0 : rec[AUXOFF + b];
            }
#pragma unroll
            for (int r = 0; r < R64; ++r) {
                double s = 1.0;
                if (MODE == FM_FS0) {
                    double inter = 0.0, shape = 0.0;
#pragma unroll
                    for (int b = 0; b < NF; ++b) {
                        inter = fma(xw[r][b], m[b], inter);
                        shape = fma(w[r][b], aux[b], shape);
                    }
                    s = inter / shape;
                }
                double chi2 = 0.0;
#pragma unroll
                for (int b = 0; b < NF; ++b) {
                    double res = fma(-s, m[b], d[r][b]);
                    double wb = (MODE == FM_FX1) ? 1.0 / (w[r][b] + aux[b]) : w[r][b];
                    chi2 = fma(res * wb, res, chi2);
                }
                const double cc = chi2 * 0.72134752044448170368;
                double l = prior2 - cc;
                if (DP) l = fma(A[r], (double)fast_lg2((float)cc), l);
                double delta = l - M[r];
                if (PASS == 1) {
                    float e = fast_ex2(-fabsf((float)delta));
                    bool gt = delta > 0.0;
                    S[r] = __fmaf_rn(S[r], gt ? e : 1.f, gt ? 1.f : e);
                    M[r] = gt ? l : M[r];
                    best[r] = gt ? (int)(first + jj) : best[r];
                    lv[r] = lv[r] || delta > P.live_lthr;
                } else {
                    float u = fast_ex2((float)delta);
                    u = (l > thr[r]) ? u : 0.f;
                    acc[r] = __fmaf_rn(u, invnorm, acc[r]);
                }
            }
        }
        if (PASS == 1) {
#pragma unroll
            for (int r = 0; r < R64; ++r) {
                Sd[r] = Sd[r] * exp2(Mfl[r] - M[r]) + (double)S[r];
                Mfl[r] = M[r];
                S[r] = 0.f;
                if (P.live && oidx[r] >= 0) P.live[(size_t)(t0 + it) * (size_t)P.No_pad + oidx[r]] = lv[r] ? 1 : 0;
            }
        }
        __syncthreads();
        if (tid == 0 && it + NSTAGE < nt) issue(it + NSTAGE);
    }
    if (PASS == 1) {
#pragma unroll
        for (int r = 0; r < R64; ++r)
            if (oidx[r] >= 0) {
                const int64_t slot = (int64_t)blockIdx.x * (FT64 * R64) + (int64_t)r * FT64 + tid;
                size_t q = P.part_stride > 0 ? (size_t)blockIdx.y * P.part_stride + slot : (size_t)blockIdx.y * P.No_pad + oidx[r];
                P.pM[q] = Mfl[r];
                P.pS[q] = Sd[r];
                P.pbest[q] = best[r];
            }
    } else if (cur_bin >= 0) {
#pragma unroll
        for (int r = 0; r < R64; ++r)
            if (acc[r] != 0.f && oidx[r] >= 0) atomicAdd(P.hist + (int64_t)oidx[r] * P.hist_stride + cur_bin, acc[r]);
    }
}

struct Rec64Params {
    const double *m, *me, *lnprior;
    const int32_t* perm;
    const int32_t* bins;
    const float* invnorm;
    int64_t nm;
    int Nf, mode, rec;
    double* recs;
};

__global__ void k_build_records64(Rec64Params P) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.nm) return;
    int64_t j = P.perm[p];
    double* r = P.recs + p * P.rec;
    for (int b = 0; b < P.Nf; ++b) {
        double v = P.m[j * P.Nf + b];
        r[b] = v;
        if (P.mode == FM_FS0) r[P.Nf + b] = v * v;
        else if (P.mode == FM_FX1) {
            double e = P.me[j * P.Nf + b];
            r[P.Nf + b] = e * e;
        }
    }
    int tail = P.Nf * ((P.mode == FM_FX0) ? 1 : 2);
    r[tail] = P.lnprior ? P.lnprior[j] * 1.4426950408889634 : 0.0;
    int2 tl;
    tl.x = P.bins ? P.bins[p] : -1;
    tl.y = __float_as_int(P.invnorm ? P.invnorm[p] : 0.f);
    *reinterpret_cast<int2*>(r + tail + 1) = tl;
    for (int i = tail + 2; i < P.rec; ++i) r[i] = 0.0;
}

// ---- object preparation ------------------------------------------------------------------------
struct PrepParams {
    const double *x, *xe, *xm;   // (No x Nf) raw inputs of this chunk
    int64_t No, No_pad;
    int Nf, mode, free_scale, dim_prior;
    float *od, *ow, *ox, *odl, *oA, *osnr;
    int32_t* obits;
    int32_t* n_not_one;   // nullable: counts the objects of the chunk whose (dof/2 - 1) != 1
};

__global__ void k_prep_objects(PrepParams P) {
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= P.No_pad) return;
    double ndim = 0.0, snr2 = 0.0;
    int bits = 0;
    bool binary = true;
    for (int b = 0; b < P.Nf; ++b) {
        double d = 0.0, e = 1.0, k = 0.0;
        if (o < P.No) {
            d = P.x[o * P.Nf + b];
            e = P.xe[o * P.Nf + b];
            k = P.xm[o * P.Nf + b];
            bool clean = isfinite(d) && isfinite(e) && (e > 0.0);   // pdf.py:310-311
            if (!clean) { d = 0.0; e = 1.0; k = 0.0; }
        }
        float dh = (float)d;
        float dl = (float)(d - (double)dh);
        double w = k / (e * e);
        size_t q = (size_t)b * P.No_pad + o;
        P.od[q] = dh;
        P.odl[q] = 2.f * dl;
        if (P.mode == FM_FX1) {
            P.ow[q] = (k != 0.0) ? (float)(e * e) : CUDART_INF_F;
            P.ox[q] = 0.f;
        } else {
            // weights carry the factor -log2(e)/2 so that the sweep accumulates c = -chi2*log2(e)/2 directly
            const double wk = -0.72134752044448170368 * w;
            P.ow[q] = (float)wk;
            P.ox[q] = (float)(d * wk);
        }
        ndim += k;
        if (k != 0.0) bits |= 1 << b;
        if (k != 0.0 && k != 1.0) binary = false;
        double sn = (k != 0.0) ? fabs(d) / e : 0.0;
        snr2 += sn * sn;
    }
    double a = P.free_scale ? 0.5 * (ndim - 1.0) : 0.5 * ndim;
    P.oA[o] = P.dim_prior ? (float)(a - 1.0) : 0.f;
    if (P.n_not_one && o < P.No && P.dim_prior && (a - 1.0) != 1.0) atomicAdd(P.n_not_one, 1);
    // a non-binary object mask makes the pair dimensionality non-integer: leave such objects to the float64 kernels
    P.osnr[o] = binary ? (float)sqrt(snr2) : CUDART_INF_F;
    if (P.obits) P.obits[o] = bits;
}

// ndim-dependent part of the ln-likelihood in sweep units (log2, chi2 scaled by log2(e)/2), used when the pair
// dimensionality varies from model to model (model masks):
//   dim_prior: -(lgamma(a) + a ln2) log2(e) - (a - 1) log2(log2(e)/2),  a = dof/2       (pdf.py:93 / :229)
//   else     : -ndim ln(2 pi)/2 log2(e)                                                  (pdf.py:96-98)
__host__ __device__ inline double fzb_ktab(double ndim, int free_scale, int dim_prior) {
    if (!dim_prior) return -0.5 * ndim * 1.83787706640934548356 * 1.4426950408889634;
    const double a = free_scale ? 0.5 * (ndim - 1.0) : 0.5 * ndim;
    return -(lgamma(a) + a * 0.69314718055994530942) * 1.4426950408889634 - (a - 1.0) * log2(0.72134752044448170368);
}

// ---- merge of the model splits + exact re-evaluation of the best pair + routing -------------------
struct MergeParams {
    const double *x, *xe, *xm;          // raw inputs (chunk)
    const double *m, *me, *mm;          // float64 models (original order)
    const double* lnprior;              // nullable
    const int32_t* perm;                // sorted position -> original model
    int64_t No, No_pad, o_base;
    int Nf, nsplit, free_scale, ime, dim_prior;
    const double* pM;
    const double* pS;
    const int32_t* pbest;
    const float* osnr;
    double log2_wt_thresh;              // log2(wt_thresh) or -inf
    double chi2_max, snr_max, consist_tol;
    int force_fp32;
    int mmv;                            // model-mask variant: the sweep value includes the ndim-dependent constant
    const float* lin_oA;                // linear-domain tensor-core sweep: objects with (dof/2 - 1) != 1 go to float64
    int stage;                          // 0: after the fp32 sweep (all objects); 1: after the float64 sweep (in_list)
    const int32_t* in_list;
    int64_t n_in;
    int64_t part_stride;                // stage 1, > 0: partials indexed by position in in_list
    // outputs
    double *lmap, *levid, *best_chi2, *best_scale;   // absolute object index
    double* Sout;                       // nullable: sum exp(l - lmap) (model-sharded pass 1)
    double* lmap_local;                 // chunk-local copy of lmap (model-sharded pass 2 needs it)
    double* Mlocal;                     // chunk-local maximum in sweep units, float64
    int64_t* best_idx;
    float *M2, *thr2;                   // chunk-local, for the fp32 pass 2
    double *M2d, *thr2d;                // chunk-local, for the float64 pass 2
    // routing: counts[0] fp32-safe (chunk-local), counts[1] degenerate -> generic kernel (absolute),
    //          counts[2] needs the float64 sweep (chunk-local), counts[3] float64-sweep objects ready for pass 2
    int32_t *safe_list, *unsafe_list, *prec_list, *safe64_list, *counts;
    // fused single pass (k_sweep_tc<..., FUSE>): an fp32-safe object keeps its histogram when every thread that worked on
    // it stayed in the frame of M0 without overflowing its record segment and the final maximum is within the recorded
    // band above M0 -> fuse_list (counts[5]); otherwise safe_list (row cleared, pruned pass 2)
    const int* fz_cnt;                  // [nsplit][No_pad] (tensor-core sweep), nullable
    const unsigned int* fz_count;       // packed sweep: records written to the CutRecord list, and its capacity
    unsigned int fz_count_cap;
    const float* fz_M0;                 // null: not a fused pass
    int fz_cap;
    double fz_glog2;                    // log2 of the band's upper edge, less a margin
    unsigned char* fz_ok;               // [No_pad] out: 1 = histogram of the fused pass stands
    int32_t* fuse_list;
};

__global__ void k_merge(MergeParams P) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t o;
    if (P.stage == 0) {
        if (i >= P.No) return;
        o = i;
    } else {
        if (i >= P.n_in) return;
        o = P.in_list[i];
    }
    double M = -DBL_MAX;
    int bs = 0;
    const bool by_pos = (P.stage == 1 && P.part_stride > 0);
    const size_t pstride = by_pos ? (size_t)P.part_stride : (size_t)P.No_pad;
    const size_t pcol = by_pos ? (size_t)i : (size_t)o;
    for (int s = 0; s < P.nsplit; ++s) {
        double v = P.pM[(size_t)s * pstride + pcol];
        // exact ties go to the smallest model position, whatever the partition of the models was
        if (v > M || (v == M && P.pbest[(size_t)s * pstride + pcol] < P.pbest[(size_t)bs * pstride + pcol])) { M = v; bs = s; }
    }
    double S = 0.0;
    bool bad = false;
    for (int s = 0; s < P.nsplit; ++s) {
        double v = P.pM[(size_t)s * pstride + pcol];
        double ss = P.pS[(size_t)s * pstride + pcol];
        if (!(ss == ss) || isinf(ss)) bad = true;
        if (v > -1e300) S += ss * exp2(v - M);
    }
    int64_t sorted_best = P.pbest[(size_t)bs * pstride + pcol];
    int64_t j = P.perm[sorted_best];
    // exact float64 evaluation of (object, best model): pdf.py:27-235 via fzb_pair64.cuh
    double sx[FZB_FAST_MAXF], sxe[FZB_FAST_MAXF], sxm[FZB_FAST_MAXF];
    for (int b = 0; b < P.Nf; ++b) {
        double d = P.x[o * P.Nf + b], e = P.xe[o * P.Nf + b], k = P.xm[o * P.Nf + b];
        bool clean = isfinite(d) && isfinite(e) && (e > 0.0);
        sx[b] = clean ? d : 0.0;
        sxe[b] = clean ? e : 1.0;
        sxm[b] = clean ? k : 0.0;
    }
    fzb64::PairState st;
    fzb64::pair_first(sx, sxe, sxm, P.m + j * P.Nf, P.me + j * P.Nf, P.mm + j * P.Nf, P.Nf, P.free_scale, P.ime, st);
    double a = P.free_scale ? 0.5 * (st.ndim - 1.0) : 0.5 * st.ndim;
    double lnl = P.dim_prior ? fzb64::chi2_logpdf(st.chi2, a) : st.lnl;
    double lp = P.lnprior ? P.lnprior[j] : 0.0;
    double lmap = P.lnprior ? lnl + lp : lnl;
    // the same "varying part" the sweeps track, in float64, for a consistency check
    const double cc = 0.72134752044448170368 * st.chi2;
    double vary = -cc + lp * 1.4426950408889634;
    if (P.dim_prior && (a - 1.0) != 0.0) vary += (a - 1.0) * log2(cc);
    if (P.mmv) vary += fzb_ktab(st.ndim, P.free_scale, P.dim_prior);
    const bool finite = !bad && isfinite(M) && M > -1e300 && isfinite(S) && S >= 0.5 && isfinite(lmap);
    const bool consistent = fabs(vary - M) <= P.consist_tol * fmax(1.0, fabs(vary));
    bool precise = consistent;
    if (P.stage == 0 && !P.force_fp32) precise = precise && (st.chi2 <= P.chi2_max) && ((double)P.osnr[o] <= P.snr_max);
    if (P.stage == 0 && P.lin_oA && P.lin_oA[o] != 1.f) precise = false;
    int64_t og = P.o_base + o;
    if (P.lmap) P.lmap[og] = lmap;
    // sum_j exp(lnprob_j) = exp(C) 2^M S with the sweep's own (fp32 or float64) maximum M, and lmap = C + vary ln 2 exactly:
    // refer the sum to the exact maximum so that the rounding error of M does not shift the evidence; the best model's own
    // term is exactly 1 then
    if (finite && consistent) S = 1.0 + fmax(S - 1.0, 0.0) * exp2(M - vary);
    if (P.levid) P.levid[og] = lmap + log(S);
    if (P.Sout) P.Sout[og] = S;
    if (P.lmap_local) P.lmap_local[o] = lmap;
    if (P.Mlocal) P.Mlocal[o] = M;
    if (P.best_idx) P.best_idx[og] = j;
    if (P.best_chi2) P.best_chi2[og] = st.chi2;
    if (P.best_scale) P.best_scale[og] = st.scale;
    if (P.stage == 0) {
        P.M2[o] = (float)M;
        P.thr2[o] = (float)(M + P.log2_wt_thresh);
        bool fok = false;
        if (P.fz_M0 && finite && precise) {
            fok = (M - (double)P.fz_M0[o]) <= P.fz_glog2;
            if (P.fz_cnt)            // tensor-core sweep: per-thread record segments, frame changes
                for (int s = 0; s < P.nsplit; ++s) {
                    const int c = P.fz_cnt[(size_t)s * P.No_pad + o];
                    if (c < 0 || c > P.fz_cap) fok = false;
                }
            if (P.fz_count && *P.fz_count > P.fz_count_cap) fok = false;      // packed sweep: one list, it must hold every record
        }
        if (P.fz_ok) P.fz_ok[o] = fok ? 1 : 0;
        if (fok) P.fuse_list[atomicAdd(&P.counts[5], 1)] = (int32_t)o;
        else if (finite && precise) P.safe_list[atomicAdd(&P.counts[0], 1)] = (int32_t)o;
        else if (finite && P.prec_list) P.prec_list[atomicAdd(&P.counts[2], 1)] = (int32_t)o;
        else P.unsafe_list[atomicAdd(&P.counts[1], 1)] = (int32_t)og;
    } else {
        P.M2d[o] = M;
        P.thr2d[o] = M + P.log2_wt_thresh;
        if (finite && precise) P.safe64_list[atomicAdd(&P.counts[3], 1)] = (int32_t)o;
        else P.unsafe_list[atomicAdd(&P.counts[1], 1)] = (int32_t)og;
    }
}

// ---- model-sharded pass 2: express the GLOBAL lmap in this shard's sweep units ------------------------
struct ShardThrParams {
    int64_t No;
    const double* g_lmap;        // global lmap (natural log, with all constants)
    const double* lmap_local;    // this shard's lmap
    const double* M2d_local;     // this shard's maximum in sweep units (float64 copy)
    double log2_wt_thresh;
    float *M2, *thr2;
    double *M2d, *thr2d;
    double* scale;               // factor applied to the fp32-path histogram of each object by k_finish
};
__global__ void k_shard_thr(ShardThrParams P) {
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= P.No) return;
    // l_sweep = (lnprob - C_o) * log2(e) with the same per-object constant C_o for every model, so
    // M_ref = M_local + (lmap_global - lmap_local) * log2(e) makes 2^(l_sweep - M_ref) = exp(lnprob - lmap_global).
    // The fp32 sweep keeps its own (exactly representable) local maximum as the reference of its weights and the
    // factor 2^(M_local - M_ref) <= 1 is applied in float64 when the histogram is convolved; the float64 sweep uses
    // M_ref directly.
    const double mloc = P.M2d_local[o];
    const double mref = mloc + (P.g_lmap[o] - P.lmap_local[o]) * 1.4426950408889634;
    P.M2[o] = (float)mloc;
    P.thr2[o] = (float)(mref + P.log2_wt_thresh);
    P.M2d[o] = mref;
    P.thr2d[o] = mref + P.log2_wt_thresh;
    P.scale[o] = exp2((double)(float)mloc - mref);
}

// ---- histogram (*) kernel, normalise, write -------------------------------------------------------
struct FinishParams {
    const float* hist;          // [No_pad][hist_stride]
    int64_t hist_stride;
    const int32_t* objlist;     // chunk-local safe objects
    int64_t o_base;
    int Ng, Ngpad, wmax, nslot;
    const int32_t* slot_sidx;   // slot -> dictionary index
    const int32_t* widths;
    const int64_t* koff;
    const double* kernels;
    double* pdfs;               // absolute rows
    float* pdfs32;              // model-sharded partial in fp32 (exchanged over NVLink), instead of `pdfs`
    int normalise;              // 0: write the un-normalised sum (model-sharded partial)
    const double* scale;        // nullable per-object factor (chunk-local), only with normalise == 0
};

__global__ void __launch_bounds__(256) k_finish(FinishParams P) {
    extern __shared__ double sm_f[];
    // histogram of the slot in shared memory, element e at e + (e >> 4): a thread reads windows that start 4 elements apart
    // (32 bytes), which on a plain layout puts 8 lanes of a warp on the same banks; the skew spreads them over all 16
    const int hlen = P.Ngpad + 8 + ((P.Ngpad + 8) >> 4) + 1;
    double* h = sm_f;                 // (Ngpad + 8) skewed: zero tail for the 4-wide windows and the padded tap groups
    double* pdf = h + hlen;           // Ng
    double* kr = pdf + P.Ng;          // 2 wmax + 8: the taps of the slot, reversed, zero padded to groups of four
    __shared__ double red[8];
    const int tid = threadIdx.x;
    const int64_t o = P.objlist[blockIdx.x];
    for (int g = tid; g < P.Ng; g += 256) pdf[g] = 0.0;
    for (int s = 0; s < P.nslot; ++s) {
        __syncthreads();
        const float* row = P.hist + o * P.hist_stride + (size_t)s * P.Ngpad;
        const int si = P.slot_sidx[s];
        const int w = P.widths[si];
        const double* kern = P.kernels + P.koff[si];
        const int nt4 = (2 * w + 1 + 3) / 4;          // tap groups
        for (int g = tid; g < P.Ngpad + 8; g += 256) h[g + (g >> 4)] = (g < P.Ngpad) ? (double)row[g] : 0.0;
        for (int i = tid; i < 4 * nt4; i += 256) kr[i] = (i <= 2 * w) ? kern[2 * w - i] : 0.0;
        __syncthreads();
        // model at grid position pos contributes kern[x - pos + w]; histogram index = pos + wmax:
        // pdf[x] = sum_i h[x + wmax - w + i] kern[2 w - i].  Four adjacent grid points per thread and four taps per step:
        // 16 DFMA on 11 shared-memory loads (the first version reloaded a tap from global memory and rotated three registers
        // per tap: 19 instructions and 8 shared-memory wavefronts per tap, 14.8 ms per 1M objects)
        for (int x0 = 4 * tid; x0 < P.Ng; x0 += 4 * 256) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            const int e0 = x0 + P.wmax - w;              // reads up to e0 + 4 nt4 + 2 < Ngpad + 8
#pragma unroll 2
            for (int g = 0; g < nt4; ++g) {
                const double k0 = kr[4 * g], k1 = kr[4 * g + 1], k2 = kr[4 * g + 2], k3 = kr[4 * g + 3];
                const int e = e0 + 4 * g;
                double v[7];
#pragma unroll
                for (int j = 0; j < 7; ++j) v[j] = h[(e + j) + ((e + j) >> 4)];
                a0 += v[0] * k0; a1 += v[1] * k0; a2 += v[2] * k0; a3 += v[3] * k0;
                a0 += v[1] * k1; a1 += v[2] * k1; a2 += v[3] * k1; a3 += v[4] * k1;
                a0 += v[2] * k2; a1 += v[3] * k2; a2 += v[4] * k2; a3 += v[5] * k2;
                a0 += v[3] * k3; a1 += v[4] * k3; a2 += v[5] * k3; a3 += v[6] * k3;
            }
            pdf[x0] += a0;
            if (x0 + 1 < P.Ng) pdf[x0 + 1] += a1;
            if (x0 + 2 < P.Ng) pdf[x0 + 2] += a2;
            if (x0 + 3 < P.Ng) pdf[x0 + 3] += a3;
        }
    }
    __syncthreads();
    double part = 0.0;
    for (int g = tid; g < P.Ng; g += 256) part += pdf[g];
    for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += red[i];
    if (!P.normalise) tot = P.scale ? 1.0 / P.scale[o] : 1.0;
    const double inv = 1.0 / tot;          // one division per object (the PDFs of this path are held to 1e-5, not to the ulp)
    if (P.pdfs32) {
        float* out32 = P.pdfs32 + (size_t)(P.o_base + o) * P.Ng;
        for (int g = tid; g < P.Ng; g += 256) out32[g] = (float)(pdf[g] * inv);
        return;
    }
    double* out = P.pdfs + (size_t)(P.o_base + o) * P.Ng;
    for (int g = tid; g < P.Ng; g += 256) out[g] = pdf[g] * inv;
}

// ---- float64 re-decision of the weights recorded at the wt_thresh cut (CutRecord, fzb_sweep_common.cuh) -------------
struct CutFixParams {
    const CutRecord* list;
    const unsigned int* count;
    unsigned int cap;
    const double *x, *xe, *xm;      // raw objects of the chunk
    const double *m, *me, *mm;      // models, original order
    const double* lnprior;          // nullable
    const int32_t* perm;            // sorted position -> original model
    const int32_t* bins;            // sorted position -> histogram bin
    const float* invnorm;           // sorted position -> 1 / kernel normalisation
    const double* lmap;             // [chunk] exact maximum of the ln-posterior (the global one in a sharded pass 2)
    double ln_wt_thresh;
    int Nf, free_scale, ime, dim_prior;
    float* hist;
    int64_t hist_stride;
    unsigned int* changed;          // statistics
    const unsigned char* ok;        // nullable: only the records of objects with ok[obj] != 0 (fused pass of the packed sweep)
};

__device__ __forceinline__ void cutfix_object(const CutFixParams& P, int obj, double* sx, double* sxe, double* sxm) {
    for (int b = 0; b < P.Nf; ++b) {
        const double d = P.x[(size_t)obj * P.Nf + b], e = P.xe[(size_t)obj * P.Nf + b], k = P.xm[(size_t)obj * P.Nf + b];
        const bool clean = isfinite(d) && isfinite(e) && (e > 0.0);
        sx[b] = clean ? d : 0.0;
        sxe[b] = clean ? e : 1.0;
        sxm[b] = clean ? k : 0.0;
    }
}
// the float64 decision for one recorded weight (pdf.py:589-591); corrects the histogram where it differs
__device__ __forceinline__ void cutfix_apply(const CutFixParams& P, int obj, int model, float weight, bool selected,
                                             const double* sx, const double* sxe, const double* sxm) {
    const int64_t j = P.perm[model];
    fzb64::PairState st;
    fzb64::pair_first(sx, sxe, sxm, P.m + j * P.Nf, P.me + j * P.Nf, P.mm + j * P.Nf, P.Nf, P.free_scale, P.ime, st);
    const double a = P.free_scale ? 0.5 * (st.ndim - 1.0) : 0.5 * st.ndim;
    double l = P.dim_prior ? fzb64::chi2_logpdf(st.chi2, a) : st.lnl;
    if (P.lnprior) l += P.lnprior[j];
    const bool sel = l > P.lmap[obj] + P.ln_wt_thresh;
    if (sel != selected) {
        const float delta = (sel ? weight : -weight) * P.invnorm[model];
        atomicAdd(P.hist + (int64_t)obj * P.hist_stride + P.bins[model], delta);
        if (P.changed) atomicAdd(P.changed, 1u);
    }
}

__global__ void k_exact_cut_fix(CutFixParams P) {
    const unsigned int n = min(*P.count, P.cap);
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const CutRecord r = P.list[i];
        if (P.ok && !P.ok[r.obj]) continue;
        double sx[FZB_FAST_MAXF], sxe[FZB_FAST_MAXF], sxm[FZB_FAST_MAXF];
        cutfix_object(P, r.obj, sx, sxe, sxm);
        cutfix_apply(P, r.obj, r.model, r.weight, r.selected != 0, sx, sxe, sxm);
    }
}

// ---- fused single pass: seed, float64 re-decision of the recorded band, clearing of the rows that take pass 2 -------
// seed of the running cut: the pre-pass maximum (two partials: the two threads of an object), lowered by a margin
__global__ void k_fuse_seed(const double* __restrict__ pM, int64_t No, int64_t No_pad, float* __restrict__ M0, int nparts) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= No_pad) return;
    float v = -FLT_MAX;
    if (o < No) {
        double m = pM[o];
        for (int s = 1; s < nparts; ++s) m = fmax(m, pM[(size_t)s * No_pad + o]);
        if (m > -1e30 && m < 1e30) v = (float)(m - 2e-4 - 1e-6 * fabs(m));
    }
    M0[o] = v;
}

// objects of a chunk by total S/N: list A (<= snr_cut: swept without the float64 remainder of the models), list B
__global__ void k_split_snr(const float* __restrict__ osnr, int64_t No, float snr_cut, int32_t* __restrict__ list_a,
                            int32_t* __restrict__ list_b, int32_t* __restrict__ counts) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = o < No;
    const bool a = valid && osnr[o] <= snr_cut;         // NaN S/N: list B
    const unsigned ma = __ballot_sync(0xffffffffu, a), mb = __ballot_sync(0xffffffffu, valid && !a);
    const int lane = threadIdx.x & 31;
    int base_a = 0, base_b = 0;
    if (lane == 0) {
        if (ma) base_a = atomicAdd(&counts[0], __popc(ma));
        if (mb) base_b = atomicAdd(&counts[1], __popc(mb));
    }
    base_a = __shfl_sync(0xffffffffu, base_a, 0);
    base_b = __shfl_sync(0xffffffffu, base_b, 0);
    const unsigned below = (1u << lane) - 1u;
    if (a) list_a[base_a + __popc(ma & below)] = (int32_t)o;
    else if (valid) list_b[base_b + __popc(mb & below)] = (int32_t)o;
}

struct FuseFixParams {
    CutFixParams C;
    const uint4* rec;        // [part][No_pad][cap] x 3
    const int* cnt;          // [part][No_pad]
    const unsigned char* ok;
    int64_t No, No_pad, nm;
    int nparts, cap;
    float mid, half;         // the band in units of the record's cut
    unsigned int* recorded;  // statistics
};
// One thread per (part, object) segment: the weights of its sub-batch records that lie in the band go to the compact
// CutRecord list (warp-aggregated append), which k_exact_cut_fix then re-decides with every lane busy.  Records beyond
// the list's capacity are only counted: the host falls back to k_fuse_fix_direct for the chunk.
__global__ void k_fuse_collect(FuseFixParams P, CutRecord* __restrict__ list, unsigned int* __restrict__ count, unsigned int cap) {
    constexpr int G = 8;            // lanes per segment: the records of a segment are read G at a time (the loads of a
                                    // thread are dependent and latency bound: one lane per segment took 10 ms per 1M objects)
    const int lane = threadIdx.x & 31, sub = lane & (G - 1);
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    int n = 0, obj = 0;
    const uint4* r = nullptr;
    if (i < P.No * P.nparts) {
        const int part = (int)(i / P.No);
        obj = (int)(i - (int64_t)part * P.No);
        if (P.ok[obj]) n = max(0, min(P.cnt[(size_t)part * P.No_pad + obj], P.cap));
        r = P.rec + ((size_t)part * P.No_pad + obj) * P.cap * 3;
    }
    const int nmax = __reduce_max_sync(0xffffffffu, n);
    for (int k = sub; k < nmax + sub; k += G) {      // warp-uniform trip count
        const bool has = k < n;
        uint4 a = make_uint4(0, 0, 0, 0), b = a, c = a;
        if (has) { a = r[3 * k]; b = r[3 * k + 1]; c = r[3 * k + 2]; }
        const int64_t first = a.x;
        const float cut = __uint_as_float(a.y);
        const float w[8] = {__uint_as_float(a.z), __uint_as_float(a.w), __uint_as_float(b.x), __uint_as_float(b.y),
                            __uint_as_float(b.z), __uint_as_float(b.w), __uint_as_float(c.x), __uint_as_float(c.y)};
        const float cmid = cut * P.mid, chalf = cut * P.half;         // the kernel's own arithmetic
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const bool hit = has && first + q < P.nm && fabsf(w[q] - cmid) <= chalf;
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m == 0u) continue;
            unsigned int base = 0;
            const int leader = __ffs(m) - 1;
            if (lane == leader) base = atomicAdd(count, (unsigned int)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (hit) {
                const unsigned int at = base + __popc(m & ((1u << lane) - 1u));
                if (at < cap) list[at] = CutRecord{obj, (int)(first + q), w[q], w[q] > cut ? 1 : 0};
            }
        }
    }
}

// the same without the list (fallback when the list would overflow)
__global__ void k_fuse_fix_direct(FuseFixParams P) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.No * P.nparts) return;
    const int part = (int)(i / P.No);
    const int obj = (int)(i - (int64_t)part * P.No);
    if (!P.ok[obj]) return;
    const int n = min(P.cnt[(size_t)part * P.No_pad + obj], P.cap);
    if (n <= 0) return;
    double sx[FZB_FAST_MAXF], sxe[FZB_FAST_MAXF], sxm[FZB_FAST_MAXF];
    cutfix_object(P.C, obj, sx, sxe, sxm);
    const uint4* r = P.rec + ((size_t)part * P.No_pad + obj) * P.cap * 3;
    for (int k = 0; k < n; ++k) {
        const uint4 a = r[3 * k], b = r[3 * k + 1], c = r[3 * k + 2];
        const int64_t first = a.x;
        const float cut = __uint_as_float(a.y);
        const float w[8] = {__uint_as_float(a.z), __uint_as_float(a.w), __uint_as_float(b.x), __uint_as_float(b.y),
                            __uint_as_float(b.z), __uint_as_float(b.w), __uint_as_float(c.x), __uint_as_float(c.y)};
        const float cmid = cut * P.mid, chalf = cut * P.half;
        for (int q = 0; q < 8; ++q) {
            if (first + q >= P.nm || !(fabsf(w[q] - cmid) <= chalf)) continue;
            cutfix_apply(P.C, obj, (int)(first + q), w[q], w[q] > cut, sx, sxe, sxm);
        }
    }
}

__global__ void k_fuse_clear_rows(const unsigned char* __restrict__ ok, float* __restrict__ hist, int64_t hist_stride) {
    const int64_t o = blockIdx.x;
    if (ok[o]) return;
    float* row = hist + o * hist_stride;
    for (int64_t g = threadIdx.x; g < hist_stride; g += blockDim.x) row[g] = 0.f;
}

// ---- record building ------------------------------------------------------------------------------
struct RecParams {
    const double *m, *me, *lnprior;
    const int32_t* perm;
    const int32_t* bins;
    const float* invnorm;
    int64_t nm;
    int Nf, mode, rec, mlo, packed, mm;
    const double* mask;     // model masks (MM)
    float* recs;
    int stride;             // record p holds the model at sorted position p * stride (coarse set: pre-pass of the fused sweep)
};

__global__ void k_build_records(RecParams P) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.nm) return;
    const int64_t ps = p * P.stride;
    int64_t j = P.perm[ps];
    float* r = P.recs + p * P.rec;
    {
        // packed layout (k_sweep2): every model value duplicated (v, v); fixed-scale modes store -m and -2*m_lo
        const int nf = P.Nf;
        const int auxoff = 2 * nf, mloff = 2 * nf * (1 + ((P.mode == FM_FX0) ? 0 : 1));
        const float sgn = (P.mode == FM_FS0) ? 1.f : -1.f;
        const int kmoff = mloff + (P.mlo ? 2 * nf : 0);
        int mbits = 0;
        for (int b = 0; b < nf; ++b) {
            double v = P.m[j * nf + b];
            if (P.mm) {
                const bool on = P.mask[j * nf + b] != 0.0;
                r[kmoff + 2 * b] = r[kmoff + 2 * b + 1] = on ? 1.f : 0.f;
                if (on) mbits |= 1 << b;
                else v = 0.0;          // masked band: m = m^2 = 0 drops it from the scale; the mask pair drops it from chi2
            }
            float hi = (float)v;
            r[2 * b] = r[2 * b + 1] = sgn * hi;
            if (P.mlo) r[mloff + 2 * b] = r[mloff + 2 * b + 1] = sgn * 2.f * (float)(v - (double)hi);
            if (P.mode == FM_FS0) r[auxoff + 2 * b] = r[auxoff + 2 * b + 1] = (float)(v * v);
            else if (P.mode == FM_FX1) {
                double e = P.me[j * nf + b];
                r[auxoff + 2 * b] = r[auxoff + 2 * b + 1] = (float)(e * e);
            }
        }
        int tail = kmoff + (P.mm ? 2 * nf : 0);
        r[tail] = r[tail + 1] = P.lnprior ? (float)(P.lnprior[j] * 1.4426950408889634) : 0.f;
        r[tail + 2] = r[tail + 3] = P.invnorm ? P.invnorm[ps] : 0.f;
        r[tail + 4] = __int_as_float(P.bins ? P.bins[ps] : -1);
        int used = tail + 5;
        if (P.mm) r[used++] = __int_as_float(mbits);
        for (int i = used; i < P.rec; ++i) r[i] = 0.f;
        return;
    }
}

// ---- host side --------------------------------------------------------------------------------------
int mode_of(const FzbConfig& cfg) {
    if (cfg.free_scale) return (cfg.ignore_model_err == 1) ? FM_FS0 : -1;
    return cfg.ignore_model_err ? FM_FX0 : FM_FX1;
}

double env_double(const char* name, double dflt) {
    const char* v = getenv(name);
    return v ? atof(v) : dflt;
}

template <int NF, int MODE, bool DP, int R, int PASS>
int launch_sweep2_mm(fzb_context* h, const SweepParams& P, dim3 grid) {
    constexpr int REC = rec2_floats(NF, MODE, true, true);
    size_t smem = (size_t)NSTAGE * TM * REC * sizeof(float) + NSTAGE * sizeof(uint64_t) + 128;
    auto kern = k_sweep2<NF, MODE, DP, true, true, R, PASS, true>;
    FZB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, ft2_of(R), smem, h->stream>>>(P);
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    return 0;
}

template <int NF, int MODE, bool DP, bool MLO, int R, int PASS, bool FUSE = false>
int launch_sweep2_t(fzb_context* h, const SweepParams& P, dim3 grid, bool prior) {
    constexpr int REC = rec2_floats(NF, MODE, MLO);
    size_t smem = (size_t)NSTAGE * TM * REC * sizeof(float) + NSTAGE * sizeof(uint64_t) + 128;
    if (prior) {
        auto kern = k_sweep2<NF, MODE, DP, MLO, true, R, PASS, false, FUSE>;
        FZB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, ft2_of(R), smem, h->stream>>>(P);
    } else {
        auto kern = k_sweep2<NF, MODE, DP, MLO, false, R, PASS, false, FUSE>;
        FZB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, ft2_of(R), smem, h->stream>>>(P);
    }
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    return 0;
}

template <int NF, int MODE, bool DP, bool MLO>
int launch_sweep_r(fzb_context* h, const SweepParams& P, dim3 grid, int R, int pass) {
    if constexpr (MLO) {
        if (P.obits != nullptr) {   // model-mask variant
            if (pass == 1) return (R == -4) ? launch_sweep2_mm<NF, MODE, DP, 4, 1>(h, P, grid)
                                            : launch_sweep2_mm<NF, MODE, DP, 2, 1>(h, P, grid);
            return (R == -4) ? launch_sweep2_mm<NF, MODE, DP, 4, 2>(h, P, grid)
                             : launch_sweep2_mm<NF, MODE, DP, 2, 2>(h, P, grid);
        }
    }
    if (pass == 3) {   // fused single pass: the default likelihood, four objects per thread
        if constexpr (MODE == FM_FX1 && DP) {
            if (R == -4 && P.obits == nullptr) return launch_sweep2_t<NF, MODE, DP, MLO, 4, 1, true>(h, P, grid, P.has_prior != 0);
        }
        fzb_set_error("fp32 path: no fused variant of this sweep");
        return 2;
    }
    if (R < 0) {   // packed kernels: R = -objects per thread
        if (pass == 1) {
            if (R == -4) return launch_sweep2_t<NF, MODE, DP, MLO, 4, 1>(h, P, grid, P.has_prior != 0);
            return launch_sweep2_t<NF, MODE, DP, MLO, 2, 1>(h, P, grid, P.has_prior != 0);
        }
        if (R == -4) return launch_sweep2_t<NF, MODE, DP, MLO, 4, 2>(h, P, grid, P.has_prior != 0);
        return launch_sweep2_t<NF, MODE, DP, MLO, 2, 2>(h, P, grid, P.has_prior != 0);
    }
    fzb_set_error("fp32 path: bad kernel selector");
    return 2;
}

template <int NF, int MODE, bool DP>
int launch_sweep_m(fzb_context* h, const SweepParams& P, dim3 grid, bool mlo, int R, int pass) {
    return mlo ? launch_sweep_r<NF, MODE, DP, true>(h, P, grid, R, pass)
               : launch_sweep_r<NF, MODE, DP, false>(h, P, grid, R, pass);
}

template <int NF>
int launch_sweep_nf(fzb_context* h, const SweepParams& P, dim3 grid, int mode, bool dp, bool mlo, int R, int pass) {
    if (mode == FM_FS0) return dp ? launch_sweep_m<NF, FM_FS0, true>(h, P, grid, mlo, R, pass)
                                  : launch_sweep_m<NF, FM_FS0, false>(h, P, grid, mlo, R, pass);
    if (mode == FM_FX0) return dp ? launch_sweep_m<NF, FM_FX0, true>(h, P, grid, mlo, R, pass)
                                  : launch_sweep_m<NF, FM_FX0, false>(h, P, grid, mlo, R, pass);
    return launch_sweep_m<NF, FM_FX1, true>(h, P, grid, mlo, R, pass);
}

template <int NF, int MODE, bool DP>
int launch_sweep64_t(fzb_context* h, const Sweep64Params& P, dim3 grid, int pass) {
    constexpr int REC = rec64_doubles(NF, MODE);
    size_t smem = (size_t)NSTAGE * TM * REC * sizeof(double) + NSTAGE * sizeof(uint64_t);
    if (pass == 1) {
        auto kern = k_sweep64<NF, MODE, DP, 1>;
        FZB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, FT64, smem, h->stream>>>(P);
    } else {
        auto kern = k_sweep64<NF, MODE, DP, 2>;
        FZB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, FT64, smem, h->stream>>>(P);
    }
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    return 0;
}

template <int NF>
int launch_sweep64_nf(fzb_context* h, const Sweep64Params& P, dim3 grid, int mode, bool dp, int pass) {
    if (mode == FM_FS0) return dp ? launch_sweep64_t<NF, FM_FS0, true>(h, P, grid, pass)
                                  : launch_sweep64_t<NF, FM_FS0, false>(h, P, grid, pass);
    if (mode == FM_FX0) return dp ? launch_sweep64_t<NF, FM_FX0, true>(h, P, grid, pass)
                                  : launch_sweep64_t<NF, FM_FX0, false>(h, P, grid, pass);
    return launch_sweep64_t<NF, FM_FX1, true>(h, P, grid, pass);
}

int launch_sweep64(fzb_context* h, const Sweep64Params& P, dim3 grid, int nf, int mode, bool dp, int pass) {
    switch (nf) {
        case 4: return launch_sweep64_nf<4>(h, P, grid, mode, dp, pass);
        case 5: return launch_sweep64_nf<5>(h, P, grid, mode, dp, pass);
        case 6: return launch_sweep64_nf<6>(h, P, grid, mode, dp, pass);
        default: break;
    }
    fzb_set_error("float64 sweep: unsupported filter count %d", nf);
    return 2;
}

int launch_sweep(fzb_context* h, const SweepParams& P, dim3 grid, int nf, int mode, bool dp, bool mlo, int R,
                 int pass) {
    switch (nf) {
        case 4: return launch_sweep_nf<4>(h, P, grid, mode, dp, mlo, R, pass);
        case 5: return launch_sweep_nf<5>(h, P, grid, mode, dp, mlo, R, pass);
        case 6: return launch_sweep_nf<6>(h, P, grid, mode, dp, mlo, R, pass);
        default: break;
    }
    fzb_set_error("fp32 path: unsupported filter count %d", nf);
    return 2;
}

}  // namespace

bool fzb_fast_supported(const fzb_context* h, const FzbConfig& cfg) {
    if (getenv("FZB_DISABLE_FAST")) return false;
    int mode = mode_of(cfg);
    if (mode < 0) return false;                                  // iterated free scale: generic path
    if (mode == FM_FX1 && !cfg.dim_prior) return false;          // per-pair sum of ln(var): generic path
    if (h->prior_nbins > 0 && h->prior_bins_n > 0) return false;   // object-conditioned prior table: generic path
    if (h->Nf < 4 || h->Nf > 6) return false;
    if (!(h->mask_all_one || h->mask_binary) || !h->models_finite) return false;   // non-binary model masks: generic path
    if (h->Nm >= (int64_t)1 << 31) return false;
    if (h->kde_mode == FZB_KDE_GRID) return false;               // exact-Gaussian KDE: generic path
    if (h->kde_mode == FZB_KDE_DICT) {
        if (!h->labels_dict_set) return false;
        if (h->labels_bad > 0) return false;                          // labels off the grid: the float64 kernel raises if selected
        if (h->Ng > 8192) return false;                              // k_finish keeps histogram + PDF in shared memory
        if (!cfg.use_wt_thresh && cfg.use_cdf_thresh) return false;   // CDF rule needs a sort: generic path
    }
    if (cfg.use_wt_thresh && !(cfg.wt_thresh > 0.0 && cfg.wt_thresh < 1.0)) return false;
    return true;
}

// Sort the models by KDE bin, build the fp32 records, the per-model kernel normalisations and the
// slot table.  Depends on (models, lnprior, dictionary, labels, likelihood mode).
static int fast_prepare_mode(fzb_context* h, int mode) {
    FastModels& F = h->fast;
    const int64_t nm = h->Nm;
    const int nf = h->Nf;
    const bool kde = (h->kde_mode == FZB_KDE_DICT && h->labels_dict_set);
    std::vector<int32_t> perm(nm);
    std::iota(perm.begin(), perm.end(), 0);
    std::vector<int32_t> bins;
    std::vector<float> invnorm;
    F.nslot = 0;
    F.slot_sidx.clear();
    int wmax = 0, Ngpad = 0;
    if (kde) {
        // slots = distinct dictionary widths in use, ascending
        std::vector<int32_t> slot_of(h->Ndict, -1);
        for (int64_t j = 0; j < nm; ++j) slot_of[h->h_ysidx[j]] = 0;
        for (int i = 0; i < h->Ndict; ++i)
            if (slot_of[i] == 0) {
                slot_of[i] = (int32_t)F.slot_sidx.size();
                F.slot_sidx.push_back(i);
                wmax = std::max(wmax, (int)h->h_widths[i]);
            }
        F.nslot = (int)F.slot_sidx.size();
        Ngpad = h->Ng + 2 * wmax;
        std::vector<int64_t> key(nm);
        for (int64_t j = 0; j < nm; ++j)
            key[j] = (int64_t)slot_of[h->h_ysidx[j]] * Ngpad + (h->h_yidx[j] + wmax);
        std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) { return key[a] < key[b]; });
        bins.resize(nm);
        invnorm.resize(nm);
        std::vector<double> kcdf((size_t)h->h_koff[h->Ndict]);
        FZB_CUDA(cudaMemcpy(kcdf.data(), h->kcdf.p, kcdf.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int64_t p = 0; p < nm; ++p) {
            int64_t j = perm[p];
            bins[p] = (int32_t)key[j];
            // edge normalisation of the truncated kernel (pdf.py:612-617)
            int64_t si = h->h_ysidx[j], pos = h->h_yidx[j], w = h->h_widths[si];
            const double* cdf = kcdf.data() + h->h_koff[si];
            int64_t low = std::max<int64_t>(pos - w, 0), high = std::min<int64_t>(pos + w + 1, h->Ng);
            int64_t lpad = low - (pos - w), hpad = high - (pos + w + 1);
            double norm = cdf[2 * w + 1 + hpad - 1];
            if (lpad != 0) norm -= cdf[lpad - 1];
            invnorm[p] = (float)(1.0 / norm);
        }
    }
    F.nf = nf;
    F.nm = nm;
    const bool mm = !h->mask_all_one;                  // model masks present (binary, checked by fzb_fast_supported)
    const bool mlo = mm || !h->models_f32_exact;      // the mask variant is only instantiated with lo parts + prior slot
    const bool packed = true;
    F.rec = rec2_floats(nf, mode, mlo, mm);
    h->fast_packed = packed;
    if (F.recs.reserve((size_t)nm * F.rec * sizeof(float) + 64) || F.perm.reserve((size_t)nm * 4 + 16)) return 1;
    FZB_CUDA(cudaMemcpyAsync(F.perm.p, perm.data(), (size_t)nm * 4, cudaMemcpyHostToDevice, h->stream));
    if (kde) {
        if (F.bins.reserve((size_t)nm * 4 + 16) || F.invnorm.reserve((size_t)nm * 4 + 16) ||
            F.d_slot_sidx.reserve((size_t)F.nslot * 4 + 16))
            return 1;
        FZB_CUDA(cudaMemcpyAsync(F.bins.p, bins.data(), (size_t)nm * 4, cudaMemcpyHostToDevice, h->stream));
        FZB_CUDA(cudaMemcpyAsync(F.invnorm.p, invnorm.data(), (size_t)nm * 4, cudaMemcpyHostToDevice, h->stream));
        FZB_CUDA(cudaMemcpyAsync(F.d_slot_sidx.p, F.slot_sidx.data(), (size_t)F.nslot * 4, cudaMemcpyHostToDevice,
                                 h->stream));
    }
    RecParams R = {};
    R.m = h->models.as<double>();
    R.me = h->models_err.as<double>();
    R.lnprior = h->has_lnprior ? h->lnprior.as<double>() : nullptr;
    R.perm = F.perm.as<int32_t>();
    R.bins = kde ? F.bins.as<int32_t>() : nullptr;
    R.invnorm = kde ? F.invnorm.as<float>() : nullptr;
    R.nm = nm; R.Nf = nf; R.mode = mode; R.rec = F.rec; R.mlo = mlo ? 1 : 0; R.packed = packed ? 1 : 0;
    R.mm = mm ? 1 : 0; R.mask = h->models_mask.as<double>();
    R.recs = F.recs.as<float>();
    R.stride = 1;
    k_build_records<<<(unsigned)((nm + 255) / 256), 256, 0, h->stream>>>(R);
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    {   // every FZB_TC_COARSE-th model of the sorted order: the pre-pass of the fused sweeps
        F.nm_coarse = (nm + FZB_TC_COARSE - 1) / FZB_TC_COARSE;
        if (F.recs_coarse.reserve((size_t)F.nm_coarse * F.rec * sizeof(float) + 64)) return 1;
        RecParams RC = R;
        RC.nm = F.nm_coarse; RC.stride = FZB_TC_COARSE; RC.recs = F.recs_coarse.as<float>();
        k_build_records<<<(unsigned)((F.nm_coarse + 255) / 256), 256, 0, h->stream>>>(RC);
        fzb_count_launch(h);
        FZB_CUDA(cudaGetLastError());
    }
    {
        Rec64Params R6 = {};
        R6.m = R.m; R6.me = R.me; R6.lnprior = R.lnprior; R6.perm = R.perm; R6.bins = R.bins; R6.invnorm = R.invnorm;
        R6.nm = nm; R6.Nf = nf; R6.mode = mode; R6.rec = rec64_doubles(nf, mode);
        if (F.recs64.reserve((size_t)nm * R6.rec * sizeof(double) + 64)) return 1;
        R6.recs = F.recs64.as<double>();
        k_build_records64<<<(unsigned)((nm + 255) / 256), 256, 0, h->stream>>>(R6);
        fzb_count_launch(h);
        FZB_CUDA(cudaGetLastError());
    }
    F.tc_valid = false;
    if (mode == FM_FS0 && !mm && nf >= 4 && nf <= 6) {
        // tensor-core sweep tiles; models that are not fp32-representable carry their float64 remainder (MLO)
        if (fzb_build_tiles_tc(h, R.lnprior, R.bins, R.invnorm, !h->models_f32_exact)) return 1;
    }
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    F.valid = true;
    h->fast_dirty = false;
    h->fast_mode = mode;
    h->fast_wmax = wmax;
    h->fast_Ngpad = Ngpad;
    return 0;
}

int fzb_fast_fit_predict_dev(fzb_context* h, const double* d_x, const double* d_xe, const double* d_xm, int64_t No,
                             const FzbConfig& cfg, double* d_pdfs, double* d_lmap, double* d_levid,
                             int64_t* d_best_idx, double* d_best_chi2, double* d_best_scale, int shard_mode,
                             double* d_psum, const double* d_glmap) {
    const int mode = mode_of(cfg);
    const int nf = h->Nf;
    const bool want_pdf = d_pdfs != nullptr || shard_mode == 1 || (shard_mode == 2 && h->shard_out32 != nullptr);   // pass 1 of a sharded run prepares the KDE layout too
    if (want_pdf) FZB_CHECK(h->kde_mode == FZB_KDE_DICT && h->labels_dict_set, "dictionary KDE not configured");
    if (shard_mode != 1) h->shard_valid = (shard_mode == 2) ? h->shard_valid : false;
    if (h->fast_dirty || !h->fast.valid || h->fast_mode != mode) {
        if (fast_prepare_mode(h, mode)) return 1;
    }
    FastModels& F = h->fast;
    const int64_t nm = F.nm;
    const bool kde = want_pdf;
    const int64_t hist_stride = kde ? (int64_t)F.nslot * h->fast_Ngpad : 0;
    FZB_CHECK(!kde || F.nslot <= 512, "too many distinct kernel widths for the fp32 path");

    // chunk the objects so that the histogram stays within ~12 GB
    int64_t chunk = No;
    if (kde) {
        int64_t fit = (int64_t)(((size_t)12 << 30) / ((size_t)hist_stride * 4));
        if (fit < 1024) fit = 1024;
        if (chunk > fit) chunk = fit;
    }
    FZB_CHECK(shard_mode == 0 || chunk == No,
              "model-sharded passes keep per-object state on the device: call them with at most %lld objects at a time",
              (long long)chunk);
    const int64_t chunk_pad = (chunk + 6143) / 6144 * 6144;
    const bool packed = h->fast_packed;
    // objects per thread: 4 for large batches; small batches use fewer so that more CTAs exist.
    // (negative R selects the packed kernel in launch_sweep_r)
    int Robj = (chunk >= 48 * 1024) ? 4 : 2;
    if (packed && getenv("FZB_FAST_R")) Robj = atoi(getenv("FZB_FAST_R")) >= 4 ? 4 : 2;
    const int R = packed ? -Robj : Robj;
    const bool mlo = !h->models_f32_exact;
    // tensor-core sweep (fzb_sweep_tc.cuh): free scale without model errors, no model masks
    const bool use_tc = F.tc_valid && mode == FM_FS0 && h->mask_all_one && getenv("FZB_NO_TC") == nullptr;
    const int64_t tile_objs = use_tc ? (int64_t)fzb_tc_tile_objects() : (int64_t)ft2_of(Robj) * Robj;
    const int64_t obj_tiles = (chunk_pad + tile_objs - 1) / tile_objs;
    const int64_t ntiles = (nm + TM - 1) / TM;
    // many short waves: small tail (the 256-object CTAs of the tensor-core sweep: 32 waves measured 1.3 % faster than 24)
    int64_t want_ctas = (int64_t)h->sm_count * (use_tc ? 32 : ((packed && Robj >= 4) ? 24 : 48));
    if (getenv("FZB_WAVES")) want_ctas = (int64_t)h->sm_count * std::max(1, atoi(getenv("FZB_WAVES")));
    int64_t nsplit = (want_ctas + obj_tiles - 1) / obj_tiles;
    if (nsplit > ntiles) nsplit = ntiles;
    if (nsplit > 256) nsplit = 256;
    if (nsplit < 1) nsplit = 1;
    const int tiles_per_split = (int)((ntiles + nsplit - 1) / nsplit);
    nsplit = (ntiles + tiles_per_split - 1) / tiles_per_split;

    // scratch: object SoA (4 x nf + 2 planes), partials, routing
    DevBuf& so = h->misc[0];
    size_t plane = (size_t)chunk_pad * sizeof(float);
    if (so.reserve(plane * (4 * nf + 2 + 2 + 1) + 256)) return 1;
    float* base = so.as<float>();
    float* od = base;
    float* ow = od + (size_t)nf * chunk_pad;
    float* ox = ow + (size_t)nf * chunk_pad;
    float* odl = ox + (size_t)nf * chunk_pad;
    float* oA = odl + (size_t)nf * chunk_pad;
    float* osnr = oA + chunk_pad;
    float* M2 = osnr + chunk_pad;
    float* thr2 = M2 + chunk_pad;
    int32_t* obits = reinterpret_cast<int32_t*>(thr2 + chunk_pad);
    const bool mm = !h->mask_all_one;
    // partial (max, sum, arg-max) per model split; the tensor-core sweep reports TC_SPLIT partials per split
    const int64_t npart = use_tc ? nsplit * fzb_tc_split() : nsplit;
    if (h->misc[1].reserve((size_t)npart * chunk_pad * 20 + 256)) return 1;
    double* pS = h->misc[1].as<double>();
    double* pM = pS + (size_t)npart * chunk_pad;
    int32_t* pbest = reinterpret_cast<int32_t*>(pM + (size_t)npart * chunk_pad);
    if (h->misc[2].reserve((size_t)chunk_pad * 16 + 64)) return 1;
    int32_t* safe_list = h->misc[2].as<int32_t>();
    int32_t* unsafe_list = safe_list + chunk_pad;
    int32_t* prec_list = unsafe_list + chunk_pad;
    int32_t* safe64_list = prec_list + chunk_pad;
    if (h->misc[3].reserve(64)) return 1;
    int32_t* counts = h->misc[3].as<int32_t>();
    if (kde && h->misc[4].reserve((size_t)chunk_pad * hist_stride * 4 + 256)) return 1;
    float* hist = kde ? h->misc[4].as<float>() : nullptr;
    // pass-2 pruning of the tensor-core sweep: live bits [model tile][half][object], sort keys of the safe list
    // (the packed sweep prunes by whole tiles: one uint16 per (tile, object) holding a single bit, so that both kernels share
    // the sort of the pass-2 list)
    const bool prune = (use_tc || packed) && !(!use_tc && !h->mask_all_one) && kde && cfg.use_wt_thresh &&
                       getenv("FZB_NO_PRUNE") == nullptr;
    const int64_t live_rows = ntiles * (use_tc ? fzb_tc_split() : 1);
    const int live_bits = use_tc ? 16 : 1;
    unsigned short* live = nullptr;
    if (prune) {
        if (h->fast.live.reserve((size_t)live_rows * chunk_pad * sizeof(unsigned short) + 256)) return 1;
        live = h->fast.live.as<unsigned short>();
        if (h->fast.sortbuf.reserve((size_t)chunk_pad * 16 + 256)) return 1;
    }
    // the float64 sweep prunes its pass 2 the same way (its objects are the bright / badly fitted ones: narrow posteriors)
    unsigned short* live64 = nullptr;
    if (kde && cfg.use_wt_thresh && getenv("FZB_NO_PRUNE") == nullptr && getenv("FZB_NO_SWEEP64") == nullptr && h->mask_all_one) {
        if (h->fast.live64.reserve((size_t)ntiles * chunk_pad * sizeof(unsigned short) + 256)) return 1;
        live64 = h->fast.live64.as<unsigned short>();
        if (h->fast.sortbuf.reserve((size_t)chunk_pad * 16 + 256)) return 1;
    }
    // weights at the wt_thresh cut are recorded by pass 2 and re-decided in float64 (k_exact_cut_fix)
    const bool exact_cut = kde && cfg.use_wt_thresh && getenv("FZB_NO_EXACT_CUT") == nullptr;
    const unsigned int cut_cap = (unsigned int)std::min<int64_t>((int64_t)1 << 27, std::max<int64_t>(1 << 16, 48 * chunk_pad));
    if (exact_cut && h->fast.cutlist.reserve((size_t)cut_cap * sizeof(CutRecord) + 64)) return 1;
    // fused single pass (tensor-core sweep, linear-domain form): coarse pre-pass + one sweep that also fills the
    // histogram; see k_sweep_tc<..., FUSE>
    const bool fuse_any = kde && cfg.use_wt_thresh && exact_cut && shard_mode == 0 && F.nm_coarse > 0 &&
                          getenv("FZB_NO_FUSE") == nullptr;
    const bool fuse_on = use_tc && fuse_any;
    // packed sweep: the reference's default likelihood (fixed scale, model errors), four objects per thread, no model masks
    const bool fuse_pk = !use_tc && packed && Robj == 4 && mode == FM_FX1 && cfg.dim_prior && h->mask_all_one && fuse_any;
    const int fz_cap = (int)std::max<int64_t>(8, std::min<int64_t>(64, 512 / npart));      // sub-batch records per thread
    const double fz_g = env_double("FZB_FUSE_BAND", 0.006);          // recorded band above the running cut, natural log units
    uint4* fz_rec = nullptr;
    int* fz_cnt = nullptr;
    float* fz_M0 = nullptr;
    unsigned char* fz_ok = nullptr;
    int32_t* fuse_list = nullptr;
    int32_t* snr_list[2] = {nullptr, nullptr};
    double *pS_c = nullptr, *pM_c = nullptr;
    int32_t* pbest_c = nullptr;
    if (fuse_on || fuse_pk) {
        const size_t n_rec = fuse_on ? (size_t)npart * chunk_pad * fz_cap * 48 : 0, n_cnt = (size_t)npart * chunk_pad * 4;
        const size_t n_c = (size_t)fzb_tc_split() * chunk_pad;
        if (F.fuse.reserve(n_rec + n_cnt + (size_t)chunk_pad * (4 + 4 + 1 + 8) + n_c * 20 + 1024)) return 1;
        fz_rec = F.fuse.as<uint4>();
        pS_c = reinterpret_cast<double*>(fz_rec + (fuse_on ? (size_t)npart * chunk_pad * fz_cap * 3 : 0));
        pM_c = pS_c + n_c;
        pbest_c = reinterpret_cast<int32_t*>(pM_c + n_c);
        fz_cnt = pbest_c + n_c;
        fz_M0 = reinterpret_cast<float*>(fz_cnt + (size_t)npart * chunk_pad);
        fuse_list = reinterpret_cast<int32_t*>(fz_M0 + chunk_pad);
        snr_list[0] = fuse_list + chunk_pad;
        snr_list[1] = snr_list[0] + chunk_pad;
        fz_ok = reinterpret_cast<unsigned char*>(snr_list[1] + chunk_pad);
    }
    // models that are not fp32-representable: only the brighter objects need the float64 remainder of the fluxes in the
    // sweep (its effect on ln L is ~2e-8 S/N); the others are swept on the fp32-rounded tile sets
    // The same cut routes the sweeps: the faint objects (broad posteriors, maximum well bounded by the pre-pass) take the
    // fused single pass, the bright ones (narrow posteriors: pass 2 prunes almost everything, the fused pass would mostly
    // be repeated) the seeded pass 1 + pruned pass 2.
    const double mlo_snr = env_double("FZB_TC_MLO_SNR", 32.0);
    const bool snr_split = fuse_on && mlo_snr > 0.0;
    if (F.aux64.reserve((size_t)chunk_pad * 40 + 64)) return 1;
    double* M2d = F.aux64.as<double>();
    double* thr2d = M2d + chunk_pad;
    double* lmap_local = thr2d + chunk_pad;
    double* M2d_local = lmap_local + chunk_pad;
    double* shard_scale = M2d_local + chunk_pad;

    const double chi2_max = env_double("FZB_FAST_CHI2_MAX", 24.0);
    // the tf32-split scale of the tensor-core sweep costs ~1e-12 S/N^2 in chi2: send the very bright objects to float64
    const double snr_max = use_tc ? env_double("FZB_TC_SNR_MAX", 8000.0) : env_double("FZB_FAST_SNR_MAX", 20000.0);
    const bool use_sweep64 = getenv("FZB_NO_SWEEP64") == nullptr && !mm;   // k_sweep64 has no model-mask variant yet

    float ms;
    for (int64_t o0 = 0; o0 < No; o0 += chunk) {
        const int64_t nc = std::min(chunk, No - o0);
        const int64_t nc_pad = chunk_pad;
        PrepParams PP = {};
        PP.x = d_x + o0 * nf; PP.xe = d_xe + o0 * nf; PP.xm = d_xm + o0 * nf;
        PP.No = nc; PP.No_pad = nc_pad; PP.Nf = nf; PP.mode = mode;
        PP.free_scale = cfg.free_scale; PP.dim_prior = cfg.dim_prior;
        PP.od = od; PP.ow = ow; PP.ox = ox; PP.odl = odl; PP.oA = oA; PP.osnr = osnr; PP.obits = obits;
        bool use_lin = use_tc && nf == 5 && cfg.dim_prior && shard_mode != 2 && getenv("FZB_NO_LIN") == nullptr;
        if (use_lin) {
            FZB_CUDA(cudaMemsetAsync(counts + 4, 0, 4, h->stream));
            PP.n_not_one = counts + 4;
        }
        k_prep_objects<<<(unsigned)((nc_pad + 255) / 256), 256, 0, h->stream>>>(PP);
        fzb_count_launch(h);
        FZB_CUDA(cudaGetLastError());
        if (use_lin) {
            // the linear-domain kernel serves (dof/2 - 1) = 1 only; the other objects take the float64 sweep, so it is
            // worth it only when they are few
            int32_t n1 = 0;
            FZB_CUDA(cudaMemcpyAsync(&n1, counts + 4, 4, cudaMemcpyDeviceToHost, h->stream));
            FZB_CUDA(cudaStreamSynchronize(h->stream));
            if ((int64_t)n1 * 50 > nc) use_lin = false;
        }
        if (shard_mode == 2) use_lin = h->shard_lin;
        else h->shard_lin = use_lin;
        h->stats.sweep_kind = use_tc ? (use_lin ? 3 : 2) : 1;

        // ---- pass 1 (fp32, every object) ------------------------------------------------------------
        int32_t hc[4] = {0, 0, 0, 0};
        Sweep64Params S6 = {};
        S6.x = PP.x; S6.xe = PP.xe; S6.xm = PP.xm; S6.No = nc; S6.No_pad = nc_pad;
        S6.free_scale = cfg.free_scale; S6.dim_prior = cfg.dim_prior;
        S6.recs = F.recs64.as<double>(); S6.nm = nm; S6.tiles_per_split = tiles_per_split;
        S6.pM = pM; S6.pS = pS; S6.pbest = pbest;
        S6.live = live64; S6.live_lthr = cfg.use_wt_thresh ? std::log2(cfg.wt_thresh) - 2e-4 : -DBL_MAX;
        SweepParams SP = {};
        SP.od = od; SP.ow = ow; SP.ox = ox; SP.odl = odl; SP.oA = oA;
        SP.obits = mm ? obits : nullptr;
        if (mm) {
            for (int n = 0; n < 16; ++n) {
                double a = cfg.free_scale ? 0.5 * (n - 1.0) : 0.5 * n;
                SP.Atab[n] = cfg.dim_prior ? (float)(a - 1.0) : 0.f;
                SP.Ktab[n] = (float)fzb_ktab((double)n, cfg.free_scale, cfg.dim_prior);
            }
        }
        SP.No_pad = nc_pad; SP.No = nc;
        SP.recs = F.recs.as<float>(); SP.nm = nm; SP.has_prior = h->has_lnprior ? 1 : 0;
        SP.tiles_per_split = tiles_per_split;
        SP.pM = pM; SP.pS = pS; SP.pbest = pbest;
        SP.live = live;
        SP.live_thr = cfg.use_wt_thresh ? (float)(cfg.wt_thresh * (1.0 - 1e-4)) : 0.f;
        SP.live_lthr = cfg.use_wt_thresh ? (float)(std::log2(cfg.wt_thresh) - 2e-4) : -FLT_MAX;
        const int64_t tiles1 = (nc + tile_objs - 1) / tile_objs;
        int64_t nsafe = 0, nsafe64 = 0, nunsafe = 0, nfuse = 0;
        unsigned int fuse_recorded = 0;
        const bool fuse_t = fuse_on && use_lin;
        const bool fuse = fuse_t || fuse_pk;
        if (shard_mode != 2) {
        FZB_CUDA(cudaEventRecord(h->ev[2], h->stream));
        if (fuse_pk) {
            // pre-pass over every FZB_TC_COARSE-th model (one split) -> seed of the running maximum
            SweepParams SC = SP;
            SC.recs = F.recs_coarse.as<float>(); SC.nm = F.nm_coarse; SC.tiles_per_split = (int)((F.nm_coarse + TM - 1) / TM);
            SC.pM = pM_c; SC.pS = pS_c; SC.pbest = pbest_c; SC.live = nullptr;
            if (launch_sweep(h, SC, dim3((unsigned)tiles1, 1u), nf, mode, cfg.dim_prior != 0, mlo, R, 1)) return 1;
            k_fuse_seed<<<(unsigned)((nc_pad + 255) / 256), 256, 0, h->stream>>>(pM_c, nc, nc_pad, fz_M0, 1);
            fzb_count_launch(h);
            FZB_CUDA(cudaGetLastError());
            FZB_CUDA(cudaMemsetAsync(hist, 0, (size_t)nc_pad * hist_stride * 4, h->stream));
            FZB_CUDA(cudaMemsetAsync(counts + 10, 0, 8, h->stream));
            h->stats.pairs_fp32 += nc * F.nm_coarse;
            // log2 domain: cut, and the recorded band from the fp32 error below it to exp(fz_g) above it
            const double l2e = 1.4426950408889634, elo = env_double("FZB_EXACT_CUT_TOL", 3e-5) * l2e, ehi = fz_g * l2e;
            SP.fz_M0 = fz_M0; SP.fz_thr = (float)std::log2(cfg.wt_thresh);
            SP.fz_mid = (float)(std::log2(cfg.wt_thresh) + 0.5 * (ehi - elo)); SP.fz_half = (float)(0.5 * (ehi + elo));
            SP.hist = hist; SP.hist_stride = hist_stride;
            SP.ex_list = h->fast.cutlist.as<CutRecord>(); SP.ex_count = reinterpret_cast<unsigned int*>(counts + 10); SP.ex_cap = cut_cap;
            if (launch_sweep(h, SP, dim3((unsigned)tiles1, (unsigned)nsplit), nf, mode, cfg.dim_prior != 0, mlo, R, 3)) return 1;
            SP.ex_list = nullptr;
        } else if (fuse) {
            // object lists of the two tile kinds (one list = all objects when the models are fp32-representable)
            int64_t nlist[2] = {0, nc};
            if (snr_split) {
                FZB_CUDA(cudaMemsetAsync(counts + 6, 0, 8, h->stream));
                k_split_snr<<<(unsigned)((nc + 255) / 256), 256, 0, h->stream>>>(osnr, nc, (float)mlo_snr, snr_list[0], snr_list[1],
                                                                                counts + 6);
                fzb_count_launch(h);
                FZB_CUDA(cudaGetLastError());
                int32_t hn[2] = {0, 0};
                FZB_CUDA(cudaMemcpyAsync(hn, counts + 6, 8, cudaMemcpyDeviceToHost, h->stream));
                FZB_CUDA(cudaStreamSynchronize(h->stream));
                nlist[0] = hn[0]; nlist[1] = hn[1];
            }
            SP.fz_M0 = fz_M0; SP.fz_thr = (float)cfg.wt_thresh;
            SP.fz_lofac = 1.f - (float)env_double("FZB_EXACT_CUT_TOL", 3e-5); SP.fz_gfac = (float)std::exp(fz_g);
            SP.fz_mid = 0.5f * (SP.fz_lofac + SP.fz_gfac); SP.fz_half = 0.5f * (SP.fz_gfac - SP.fz_lofac);
            SP.fz_rec = fz_rec; SP.fz_cnt = fz_cnt; SP.fz_cap = fz_cap;
            SP.hist = hist; SP.hist_stride = hist_stride;
            // pre-pass over every FZB_TC_COARSE-th model -> lower bound of every object's maximum
            for (int v = 0; v < 2; ++v) {
                if (nlist[v] == 0) continue;
                const bool vm = (v == 1) ? mlo : false;
                SweepParams SC = SP;
                SC.nm = F.nm_coarse; SC.tiles_per_split = (int)((F.nm_coarse + TM - 1) / TM);
                SC.pM = pM_c; SC.pS = pS_c; SC.pbest = pbest_c; SC.live = nullptr;
                if (snr_split) { SC.objlist = snr_list[v]; SC.No = nlist[v]; }
                const unsigned char* tl = (v == 1 || !mlo) ? F.tiles_tc_coarse.as<unsigned char>() : F.tiles_tc_f32_coarse.as<unsigned char>();
                if (fzb_launch_sweep_tc(h, SC, dim3((unsigned)((nlist[v] + tile_objs - 1) / tile_objs), 1u), nf, true, 1, true, vm, tl))
                    return 1;
            }
            k_fuse_seed<<<(unsigned)((nc_pad + 255) / 256), 256, 0, h->stream>>>(pM_c, nc, nc_pad, fz_M0, fzb_tc_split());
            fzb_count_launch(h);
            FZB_CUDA(cudaGetLastError());
            FZB_CUDA(cudaMemsetAsync(hist, 0, (size_t)nc_pad * hist_stride * 4, h->stream));
            h->stats.pairs_fp32 += nc * F.nm_coarse;
            for (int v = 0; v < 2; ++v) {
                if (nlist[v] == 0) continue;
                const bool vm = (v == 1) ? mlo : false;
                SweepParams SF = SP;
                if (snr_split) { SF.objlist = snr_list[v]; SF.No = nlist[v]; }
                const unsigned char* tl = (v == 1 || !mlo) ? nullptr : F.tiles_tc_f32.as<unsigned char>();
                const int variant = (snr_split && v == 1 && getenv("FZB_FUSE_BRIGHT") == nullptr) ? 2 : 1;
                if (fzb_launch_sweep_tc(h, SF, dim3((unsigned)((nlist[v] + tile_objs - 1) / tile_objs), (unsigned)nsplit), nf, true, 1,
                                        true, vm, tl, variant))
                    return 1;
            }
        } else if (use_tc ? fzb_launch_sweep_tc(h, SP, dim3((unsigned)tiles1, (unsigned)nsplit), nf, cfg.dim_prior != 0, 1, use_lin, mlo)
                          : launch_sweep(h, SP, dim3((unsigned)tiles1, (unsigned)nsplit), nf, mode, cfg.dim_prior != 0, mlo, R, 1))
            return 1;
        FZB_CUDA(cudaEventRecord(h->ev[3], h->stream));
        h->stats.pairs_fp32 += nc * nm;

        FZB_CUDA(cudaMemsetAsync(counts, 0, 32, h->stream));
        MergeParams MP = {};
        MP.x = PP.x; MP.xe = PP.xe; MP.xm = PP.xm;
        MP.m = h->models.as<double>(); MP.me = h->models_err.as<double>(); MP.mm = h->models_mask.as<double>();
        MP.lnprior = h->has_lnprior ? h->lnprior.as<double>() : nullptr;
        MP.perm = F.perm.as<int32_t>();
        MP.No = nc; MP.No_pad = nc_pad; MP.o_base = o0;
        MP.Nf = nf; MP.nsplit = (int)npart; MP.free_scale = cfg.free_scale; MP.ime = cfg.ignore_model_err != 0;
        MP.dim_prior = cfg.dim_prior;
        MP.pM = pM; MP.pS = pS; MP.pbest = pbest; MP.osnr = osnr;
        MP.log2_wt_thresh = cfg.use_wt_thresh ? std::log2(cfg.wt_thresh) : -INFINITY;
        MP.chi2_max = chi2_max; MP.snr_max = snr_max; MP.consist_tol = 1e-4;
        MP.force_fp32 = (cfg.precision == FZB_PREC_FP32);
        MP.mmv = mm ? 1 : 0;
        MP.lin_oA = use_lin ? oA : nullptr;
        MP.stage = 0;
        MP.lmap = d_lmap; MP.levid = d_levid; MP.best_chi2 = d_best_chi2; MP.best_scale = d_best_scale;
        MP.best_idx = d_best_idx;
        MP.Sout = (shard_mode == 1) ? d_psum : nullptr;
        MP.lmap_local = lmap_local; MP.Mlocal = M2d_local;
        MP.M2 = M2; MP.thr2 = thr2; MP.M2d = M2d; MP.thr2d = thr2d;
        MP.safe_list = safe_list; MP.unsafe_list = unsafe_list; MP.prec_list = use_sweep64 ? prec_list : nullptr;
        MP.safe64_list = safe64_list; MP.counts = counts;
        if (fuse) {
            MP.fz_cnt = fuse_t ? fz_cnt : nullptr; MP.fz_M0 = fz_M0; MP.fz_cap = fz_cap; MP.fz_ok = fz_ok; MP.fuse_list = fuse_list;
            if (fuse_pk) { MP.fz_count = reinterpret_cast<unsigned int*>(counts + 10); MP.fz_count_cap = cut_cap; }
            MP.fz_glog2 = (fz_g - 1e-3) * 1.4426950408889634 - 4e-4;     // the seed's margin and the fp32 error of the weights
        }
        k_merge<<<(unsigned)((nc + 255) / 256), 256, 0, h->stream>>>(MP);
        fzb_count_launch(h);
        FZB_CUDA(cudaGetLastError());
        int32_t hc8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        FZB_CUDA(cudaMemcpyAsync(hc8, counts, 32, cudaMemcpyDeviceToHost, h->stream));
        FZB_CUDA(cudaStreamSynchronize(h->stream));
        for (int i = 0; i < 4; ++i) hc[i] = hc8[i];
        FZB_CUDA(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]));
        h->stats.ms_scan += ms;
        nsafe = hc[0];
        nfuse = fuse ? hc8[5] : 0;
        h->stats.objects_fused += nfuse;
        MP.fz_cnt = nullptr; MP.fz_ok = nullptr; MP.fz_M0 = nullptr;      // stage 1 (float64 sweep) routes as before
        const int64_t nprec = hc[2];

        // ---- objects whose fp32 result is not trusted: float64 sweep, pass 1 ---------------------------
        if (nprec > 0) {
            // few objects, all models: split the models finely enough to fill the GPU (partials compact, by list position)
            S6.objlist = prec_list; S6.nlist = nprec;
            const int64_t t64 = (nprec + FT64 * R64 - 1) / (FT64 * R64);
            int64_t ns64 = ((int64_t)h->sm_count * 4 + t64 - 1) / t64;
            ns64 = std::min(ns64, std::max<int64_t>(1, ((int64_t)32 << 20) / (t64 * FT64 * R64)));
            ns64 = std::max<int64_t>(1, std::min(ns64, ntiles));
            const int tps64 = (int)((ntiles + ns64 - 1) / ns64);
            ns64 = (ntiles + tps64 - 1) / tps64;
            const int64_t stride64 = t64 * FT64 * R64;
            if (h->misc[6].reserve((size_t)ns64 * stride64 * 20 + 256)) return 1;
            S6.pS = h->misc[6].as<double>();
            S6.pM = S6.pS + (size_t)ns64 * stride64;
            S6.pbest = reinterpret_cast<int32_t*>(S6.pM + (size_t)ns64 * stride64);
            S6.part_stride = stride64;
            S6.tiles_per_split = tps64;
            h->fast_ns64 = (int)ns64;
            if (launch_sweep64(h, S6, dim3((unsigned)t64, (unsigned)ns64), nf, mode, cfg.dim_prior != 0, 1)) return 1;
            h->stats.pairs_fp64 += nprec * nm;
            MP.stage = 1; MP.in_list = prec_list; MP.n_in = nprec; MP.consist_tol = 1e-6; MP.nsplit = (int)ns64;
            MP.pM = S6.pM; MP.pS = S6.pS; MP.pbest = S6.pbest; MP.part_stride = stride64;
            k_merge<<<(unsigned)((nprec + 255) / 256), 256, 0, h->stream>>>(MP);
            fzb_count_launch(h);
            FZB_CUDA(cudaGetLastError());
            FZB_CUDA(cudaMemcpyAsync(hc, counts, 16, cudaMemcpyDeviceToHost, h->stream));
            FZB_CUDA(cudaStreamSynchronize(h->stream));
            nsafe64 = hc[3];
        }
        nunsafe = hc[1];
        h->stats.objects_fp64 += nunsafe + nprec;
        if (shard_mode == 1) {
            // model-sharded pass 1: partials are out; remember the routing for pass 2
            if (nunsafe > 0 && fzb_generic_shard_pass1_dev(h, d_x, d_xe, d_xm, No, unsafe_list, nunsafe, cfg, d_lmap,
                                                           d_psum, d_best_idx))
                return 1;
            if (prune && nsafe > 0 && fzb_sort_by_live_bits(h, live, live_rows, nc_pad, safe_list, nsafe, live_bits))
                return 1;
            if (live64 && nsafe64 > 0 && fzb_sort_by_live_bits(h, live64, ntiles, nc_pad, safe64_list, nsafe64, 1)) return 1;
            h->shard_valid = true;
            h->shard_No = No;
            h->shard_counts[0] = (int)nsafe; h->shard_counts[1] = (int)nunsafe; h->shard_counts[3] = (int)nsafe64;
            continue;
        }
        } else {
            // model-sharded pass 2: routing lists and local maxima are still in the context; move the reference
            // of the weights and of the selection cut to the GLOBAL lmap
            FZB_CHECK(h->shard_valid && h->shard_No == No, "fzb_shard_pass2_dev without a matching pass 1");
            nsafe = h->shard_counts[0]; nunsafe = h->shard_counts[1]; nsafe64 = h->shard_counts[3];
            ShardThrParams TP = {};
            TP.No = nc; TP.g_lmap = d_glmap; TP.lmap_local = lmap_local; TP.M2d_local = M2d_local;
            TP.log2_wt_thresh = cfg.use_wt_thresh ? std::log2(cfg.wt_thresh) : -INFINITY;
            TP.M2 = M2; TP.thr2 = thr2; TP.M2d = M2d; TP.thr2d = thr2d; TP.scale = shard_scale;
            k_shard_thr<<<(unsigned)((nc + 255) / 256), 256, 0, h->stream>>>(TP);
            fzb_count_launch(h);
            FZB_CUDA(cudaGetLastError());
        }

        if (nunsafe > 0) {
            // degenerate rows (non-finite sums, NaN-producing models, ...): generic float64 kernel with the
            // reference's exact semantics; objsel holds absolute object indices
            if (shard_mode == 2) {
                if (fzb_generic_shard_pass2_dev(h, d_x, d_xe, d_xm, No, unsafe_list, nunsafe, cfg, d_glmap, d_glmap, d_pdfs))
                    return 1;
            } else if (fzb_generic_fit_predict_dev(h, d_x, d_xe, d_xm, No, unsafe_list, nunsafe, cfg, d_pdfs, d_lmap,
                                                   d_levid, d_best_idx, d_best_chi2, d_best_scale))
                return 1;
        }
        if (kde && (nsafe > 0 || nsafe64 > 0 || nfuse > 0)) {
            FZB_CUDA(cudaEventRecord(h->ev[4], h->stream));
            if (fuse) {      // rows of the objects that did not keep their fused histogram start from zero again
                k_fuse_clear_rows<<<(unsigned)nc, 256, 0, h->stream>>>(fz_ok, hist, hist_stride);
                fzb_count_launch(h);
                FZB_CUDA(cudaGetLastError());
                if (fuse_pk && nfuse > 0) {
                    // the band records of the fused pass (objects that keep their histogram only), before pass 2 reuses the list
                    CutFixParams CA = {};
                    CA.list = h->fast.cutlist.as<CutRecord>(); CA.count = reinterpret_cast<unsigned int*>(counts + 10);
                    CA.cap = cut_cap; CA.x = PP.x; CA.xe = PP.xe; CA.xm = PP.xm;
                    CA.m = h->models.as<double>(); CA.me = h->models_err.as<double>(); CA.mm = h->models_mask.as<double>();
                    CA.lnprior = h->has_lnprior ? h->lnprior.as<double>() : nullptr;
                    CA.perm = F.perm.as<int32_t>(); CA.bins = F.bins.as<int32_t>(); CA.invnorm = F.invnorm.as<float>();
                    CA.lmap = lmap_local; CA.ln_wt_thresh = std::log(cfg.wt_thresh);
                    CA.Nf = nf; CA.free_scale = cfg.free_scale; CA.ime = cfg.ignore_model_err != 0; CA.dim_prior = cfg.dim_prior;
                    CA.hist = hist; CA.hist_stride = hist_stride; CA.changed = reinterpret_cast<unsigned int*>(counts + 13);
                    CA.ok = fz_ok;
                    FZB_CUDA(cudaMemsetAsync(counts + 13, 0, 4, h->stream));
                    k_exact_cut_fix<<<h->sm_count * 8, 256, 0, h->stream>>>(CA);
                    fzb_count_launch(h);
                    FZB_CUDA(cudaGetLastError());
                    unsigned int na[4] = {0, 0, 0, 0};
                    FZB_CUDA(cudaMemcpyAsync(na, counts + 10, 16, cudaMemcpyDeviceToHost, h->stream));
                    FZB_CUDA(cudaStreamSynchronize(h->stream));
                    fuse_recorded = na[0];
                    h->stats.cut_changed += na[3];
                    FZB_CUDA(cudaMemsetAsync(counts + 10, 0, 8, h->stream));      // pass 2 (if any) starts its own list
                }
            } else {
                FZB_CUDA(cudaMemsetAsync(hist, 0, (size_t)nc_pad * hist_stride * 4, h->stream));
            }
            if (nsafe > 0) {
                if (prune && shard_mode != 2 &&
                    fzb_sort_by_live_bits(h, live, live_rows, nc_pad, safe_list, nsafe, live_bits))
                    return 1;       // objects with similar survivor sets share a warp (sharded pass 2: sorted by pass 1)
                if (prune) {
                    FZB_CUDA(cudaMemsetAsync(counts + 8, 0, 8, h->stream));
                    SP.pairs_done = reinterpret_cast<unsigned long long*>(counts + 8);
                    unsigned int* tm = nullptr;      // chunks / tiles no object of an M-tile needs: no MMA, no TMA
                    if (use_tc && getenv("FZB_NO_TILE_SKIP") == nullptr) {
                        if (fzb_tile_masks(h, live, ntiles, nc_pad, safe_list, nsafe, (int)tile_objs, &tm)) return 1;
                    }
                    SP.tmask = tm;
                }
                if (exact_cut) {
                    FZB_CUDA(cudaMemsetAsync(counts + 10, 0, 8, h->stream));
                    SP.ex_list = h->fast.cutlist.as<CutRecord>();
                    SP.ex_count = reinterpret_cast<unsigned int*>(counts + 10);
                    SP.ex_cap = cut_cap;
                    SP.ex_tol = (float)env_double("FZB_EXACT_CUT_TOL", 3e-5);
                }
                SP.No = nsafe; SP.objlist = safe_list; SP.M2 = M2; SP.thr2 = thr2; SP.hist = hist;
                SP.hist_stride = hist_stride;
                const int64_t tiles2 = (nsafe + tile_objs - 1) / tile_objs;
                if (use_tc ? fzb_launch_sweep_tc(h, SP, dim3((unsigned)tiles2, (unsigned)nsplit), nf, cfg.dim_prior != 0, 2, use_lin, mlo)
                           : launch_sweep(h, SP, dim3((unsigned)tiles2, (unsigned)nsplit), nf, mode, cfg.dim_prior != 0, mlo, R, 2))
                    return 1;
                h->stats.pairs_fp32 += nsafe * nm;
            }
            if (nsafe64 > 0) {
                if (live64 && shard_mode != 2 && fzb_sort_by_live_bits(h, live64, ntiles, nc_pad, safe64_list, nsafe64, 1)) return 1;
                S6.objlist = safe64_list; S6.nlist = nsafe64; S6.M2 = M2d; S6.thr2 = thr2d; S6.hist = hist;
                S6.hist_stride = hist_stride;
                const int64_t t64 = (nsafe64 + FT64 * R64 - 1) / (FT64 * R64);
                int64_t ns64 = std::max<int64_t>(1, std::min<int64_t>(((int64_t)h->sm_count * 4 + t64 - 1) / t64, ntiles));
                const int tps64 = (int)((ntiles + ns64 - 1) / ns64);
                ns64 = (ntiles + tps64 - 1) / tps64;
                S6.tiles_per_split = tps64;
                if (launch_sweep64(h, S6, dim3((unsigned)t64, (unsigned)ns64), nf, mode, cfg.dim_prior != 0, 2)) return 1;
                h->stats.pairs_fp64 += nsafe64 * nm;
            }
            if (exact_cut && nsafe > 0) {
                CutFixParams CF = {};
                CF.list = h->fast.cutlist.as<CutRecord>(); CF.count = reinterpret_cast<unsigned int*>(counts + 10);
                CF.cap = cut_cap; CF.x = PP.x; CF.xe = PP.xe; CF.xm = PP.xm;
                CF.m = h->models.as<double>(); CF.me = h->models_err.as<double>(); CF.mm = h->models_mask.as<double>();
                CF.lnprior = h->has_lnprior ? h->lnprior.as<double>() : nullptr;
                CF.perm = F.perm.as<int32_t>(); CF.bins = F.bins.as<int32_t>(); CF.invnorm = F.invnorm.as<float>();
                CF.lmap = (shard_mode == 2) ? d_glmap : lmap_local;
                CF.ln_wt_thresh = std::log(cfg.wt_thresh);
                CF.Nf = nf; CF.free_scale = cfg.free_scale; CF.ime = cfg.ignore_model_err != 0; CF.dim_prior = cfg.dim_prior;
                CF.hist = hist; CF.hist_stride = hist_stride; CF.changed = reinterpret_cast<unsigned int*>(counts + 11);
                k_exact_cut_fix<<<h->sm_count * 2, 256, 0, h->stream>>>(CF);
                fzb_count_launch(h);
                FZB_CUDA(cudaGetLastError());
            }
            if (nfuse > 0 && fuse_t) {
                FuseFixParams FF = {};
                CutFixParams& CF = FF.C;
                CF.x = PP.x; CF.xe = PP.xe; CF.xm = PP.xm;
                CF.m = h->models.as<double>(); CF.me = h->models_err.as<double>(); CF.mm = h->models_mask.as<double>();
                CF.lnprior = h->has_lnprior ? h->lnprior.as<double>() : nullptr;
                CF.perm = F.perm.as<int32_t>(); CF.bins = F.bins.as<int32_t>(); CF.invnorm = F.invnorm.as<float>();
                CF.lmap = lmap_local;
                CF.ln_wt_thresh = std::log(cfg.wt_thresh);
                CF.Nf = nf; CF.free_scale = cfg.free_scale; CF.ime = cfg.ignore_model_err != 0; CF.dim_prior = cfg.dim_prior;
                CF.hist = hist; CF.hist_stride = hist_stride; CF.changed = reinterpret_cast<unsigned int*>(counts + 11);
                FF.rec = fz_rec; FF.cnt = fz_cnt; FF.ok = fz_ok; FF.No = nc; FF.No_pad = nc_pad; FF.nparts = (int)npart; FF.nm = nm;
                FF.cap = fz_cap; FF.recorded = nullptr;
                FF.mid = SP.fz_mid; FF.half = SP.fz_half;
                if (!(exact_cut && nsafe > 0)) FZB_CUDA(cudaMemsetAsync(counts + 10, 0, 8, h->stream));
                // in-band weights -> compact list (after pass 2's own records have been consumed) -> float64 re-decision
                unsigned int* fcount = reinterpret_cast<unsigned int*>(counts + 12);
                FZB_CUDA(cudaMemsetAsync(fcount, 0, 4, h->stream));
                k_fuse_collect<<<(unsigned)((nc * npart * 8 + 255) / 256), 256, 0, h->stream>>>(FF, h->fast.cutlist.as<CutRecord>(), fcount,
                                                                                       cut_cap);
                fzb_count_launch(h);
                FZB_CUDA(cudaGetLastError());
                unsigned int nrec = 0;
                FZB_CUDA(cudaMemcpyAsync(&nrec, fcount, 4, cudaMemcpyDeviceToHost, h->stream));
                FZB_CUDA(cudaStreamSynchronize(h->stream));
                fuse_recorded = nrec;
                if (nrec <= cut_cap) {
                    CutFixParams CL = FF.C;
                    CL.list = h->fast.cutlist.as<CutRecord>(); CL.count = fcount; CL.cap = cut_cap;
                    k_exact_cut_fix<<<h->sm_count * 8, 256, 0, h->stream>>>(CL);
                } else {
                    k_fuse_fix_direct<<<(unsigned)((nc * npart + 127) / 128), 128, 0, h->stream>>>(FF);
                }
                fzb_count_launch(h);
                FZB_CUDA(cudaGetLastError());
            }
            FZB_CUDA(cudaEventRecord(h->ev[5], h->stream));
            FinishParams FP = {};
            FP.hist = hist; FP.hist_stride = hist_stride; FP.o_base = o0;
            FP.Ng = h->Ng; FP.Ngpad = h->fast_Ngpad; FP.wmax = h->fast_wmax; FP.nslot = F.nslot;
            FP.slot_sidx = F.d_slot_sidx.as<int32_t>(); FP.widths = h->widths.as<int32_t>();
            FP.koff = h->koff.as<int64_t>(); FP.kernels = h->kernels.as<double>(); FP.pdfs = d_pdfs;
            FP.pdfs32 = (shard_mode == 2) ? h->shard_out32 : nullptr;
            FP.normalise = (shard_mode == 2) ? 0 : 1;
            size_t smem = sizeof(double) * ((size_t)h->fast_Ngpad + 8 + (((size_t)h->fast_Ngpad + 8) >> 4) + 1 + h->Ng +
                                            2 * (size_t)h->fast_wmax + 8);
            FZB_CUDA(cudaFuncSetAttribute(k_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (nsafe > 0) {
                FP.objlist = safe_list;
                FP.scale = (shard_mode == 2) ? shard_scale : nullptr;
                k_finish<<<(unsigned)nsafe, 256, smem, h->stream>>>(FP);
                fzb_count_launch(h);
            }
            if (nfuse > 0) {
                FP.objlist = fuse_list;
                FP.scale = nullptr;
                k_finish<<<(unsigned)nfuse, 256, smem, h->stream>>>(FP);
                fzb_count_launch(h);
            }
            if (nsafe64 > 0) {
                FP.objlist = safe64_list;
                FP.scale = nullptr;
                k_finish<<<(unsigned)nsafe64, 256, smem, h->stream>>>(FP);
                fzb_count_launch(h);
            }
            FZB_CUDA(cudaGetLastError());
            FZB_CUDA(cudaEventRecord(h->ev[6], h->stream));
            unsigned long long pdone = 0;
            unsigned int cutc[2] = {0, 0};
            if (prune && nsafe > 0)
                FZB_CUDA(cudaMemcpyAsync(&pdone, counts + 8, 8, cudaMemcpyDeviceToHost, h->stream));
            if (exact_cut && (nsafe > 0 || nfuse > 0))
                FZB_CUDA(cudaMemcpyAsync(cutc, counts + 10, 8, cudaMemcpyDeviceToHost, h->stream));
            FZB_CUDA(cudaStreamSynchronize(h->stream));
            h->stats.cut_recorded += cutc[0] + fuse_recorded;
            h->stats.cut_changed += cutc[1];
            h->stats.pairs_pass2 += (prune && nsafe > 0) ? (int64_t)pdone : nsafe * nm;
            h->stats.pairs_pass2 += nsafe64 * nm;
            FZB_CUDA(cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]));
            h->stats.ms_accum += ms;
            FZB_CUDA(cudaEventElapsedTime(&ms, h->ev[5], h->ev[6]));
            h->stats.ms_finish += ms;
        }
    }
    return 0;
}
