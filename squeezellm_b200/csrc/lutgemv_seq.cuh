// lutgemv_seq.cuh - the SEQUENCE kernel: one persistent launch runs a whole list of dependent LUT-GEMVs (a decode token's QuantLinear
// matvecs: llama.py:226-234 calls them one by one, 224 launches per LLaMA-7B token).  Included by lutgemv_kernels.cu after lutgemv_v2.cuh;
// the per-position math, table builders, outlier warps and mailboxes are v2's (namespace v2), what is new is everything BETWEEN two GEMVs.
//
// Why: with one launch per GEMV (v2 + PDL + CUDA graph) a 4096x4096 layer took 8.9 us in the chain for 1.5 us of HBM time; a fit over the
// four 7B shapes gave ~6 us of fixed cost per launch (CTA exit -> grid completion -> dependency release -> CTA start, barrier init, tensor
// map fetch, x staging, ring fill from cold) - 128 launches per token, 40 % of the step (profiles/r02_bench_llama7b_w4_s45.json).  Here
//   * grid = #SMs, launched ONCE per token; every CTA walks the same list of GEMV descriptors (global memory, written at create time);
//   * the TMA producer warp never stops at a GEMV boundary: while the consumers finish GEMV g and wait for its result, the ring already
//     holds the first 128 KB per SM of GEMV g+1 - weights do not depend on activations;
//   * the dependency itself is data flow, not a grid-wide event: y is written as SELF-VALIDATING 32-bit words {fp16 value, 16-bit tag of
//     the token} (single-copy atomic: no fence, no flag, no second round trip), the consumers of the next GEMV poll the words of their x
//     straight out of L2 and repack them into fp16 in shared memory.  A stale word carries the previous token's tag, so nothing is ever
//     reset.  The cross-CTA mailboxes of v2 carry a (token, GEMV) tag the same way.
//   * on several GPUs (column shards) the owner of a strip stores its tagged words into EVERY rank's arena over NVLink; the poll of the
//     next GEMV is the whole exchange - no collective, no flag round trip (v2's exchange waited ~10 us per launch on a system-scope flag).
#pragma once

#ifndef SQLLM_SEQ_POLL_NS
#define SQLLM_SEQ_POLL_NS 0   // pause between two reads of an x word that is not there yet
#endif

#ifndef SQLLM_SEQ_PREFETCH
#define SQLLM_SEQ_PREFETCH 0  // 1: consumer loop fetches the next stage's words before the current stage's gathers (register rotation); measured: no gain (4-bit), spills (3-bit)
#endif

namespace seq {
using namespace v2;

struct alignas(128) SeqDesc {
    CUtensorMap tm_big, tm_small;   // 64 columns x (32 units | 2 units) boxes over the packed matrix
    P2 p;                           // the GEMV: buffers, stream-K plan, workspace pointers (accumulator / flags alternate by GEMV parity)
    const uint32_t *x_tag;          // first tagged word of x in the local arena; null: p.x is a plain fp16 vector (the token's input)
    unsigned long long y_off;       // byte offset of this GEMV's tagged output vector in every rank's arena
};
struct SeqExport {
    const uint32_t *src;            // tagged words in the local arena
    void *dst;                      // plain fp16
    int n;
};
struct SeqCfg {
    const SeqDesc *descs;
    int ngemv;
    const unsigned *epoch;          // token counter (bumped by a one-thread kernel before every launch)
    int *err;                       // a bounded wait gave up
    unsigned smem_raw;
    int xbytes, nstage;
    int world, rank;
    const unsigned long long *peer_base;  // arena base address of every rank (world > 1)
    unsigned long long arena_base;        // ... and our own
    const SeqExport *exports;
    int nexport;
};

__global__ void seq_bump_epoch(unsigned *epoch) { *epoch = *epoch + 1u; }

__device__ __forceinline__ void bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
template <bool SYS>
__device__ __forceinline__ uint4 ld_relaxed_v4(const uint32_t *p) {
    uint4 v;
    if constexpr (SYS) asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool tags_ok(const uint4 w, const uint32_t want) {
    return (((w.x ^ want) | (w.y ^ want) | (w.z ^ want) | (w.w ^ want)) >> 16) == 0u;
}
// Four tagged words (16 bytes) at a time, up to four pieces per thread in flight; a piece that is not there yet is simply read again
// (bounded: 2 s, then the error word).  Returns through `sink(piece index, lo, hi)`: two packed half2.
template <bool SYS, typename F>
__device__ __forceinline__ void poll_tagged(const uint32_t *src, const int n4, const int tid, const int nthr, const uint32_t want, int *err, F sink) {
    for (int i0 = tid; i0 < n4; i0 += 4 * nthr) {
        uint4 w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i0 + j * nthr < n4) w[j] = ld_relaxed_v4<SYS>(src + 4 * (size_t)(i0 + j * nthr));
        unsigned long long t0 = 0ull, t1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = i0 + j * nthr;
            if (i < n4) {
                while (!tags_ok(w[j], want)) {
#if SQLLM_SEQ_POLL_NS > 0
                    __nanosleep(SQLLM_SEQ_POLL_NS);
#endif
                    w[j] = ld_relaxed_v4<SYS>(src + 4 * (size_t)i);
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t0 == 0ull) t0 = t1;
                    if (t1 - t0 > 2000000000ull) { *reinterpret_cast<volatile int *>(err) = 1; break; }
                }
                sink(i, __byte_perm(w[j].x, w[j].y, 0x5410), __byte_perm(w[j].z, w[j].w, 0x5410));
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------------
// Outlier warps of the sequence kernel.  v2::sparse2 stages the (col, val) chunks in shared memory, forms the products in place and sums
// rows from there: four or more DEPENDENT shared-memory round trips per chunk, each queueing for hundreds of cycles behind the consumers'
// gathers (the LSU is what bounds the kernel) - the timeline of a 4096x4096 layer showed the outlier warps finishing 7 us after x arrived
// for 510 non-zeros per CTA, 2.3 us after the last consumer warp: they, not the weights, set the layer's critical path
// (profiles/r02_seq_trace_1.txt).  Here nothing but x goes through shared memory:
//   * LPR lanes share a row (LPR from the average row length), 32 / LPR rows per pass; a lane keeps its <= 8 (col, val) pairs of the
//     pass in REGISTERS, loaded straight from global memory (L2: prefetched at entry) - everything static, so the first pass is in
//     registers before x arrives and the loads of pass k+1 are issued before pass k is summed;
//   * after x: one round of x gathers (the only shared-memory accesses), FMAs, a shuffle tree over the LPR lanes, and ONE 64-bit tagged
//     store of the row sum into the row's mailbox word (rows without outliers store 0: the strip's owner waits for all its rows);
//   * the part of a row beyond 8 x LPR elements is summed by the whole warp (skewed rows).
// Dense rows (topX): as in v2 (the CTA's k-slice, requested before x arrives, red.add + one announcement per strip).
// ---------------------------------------------------------------------------------------------------------------------------
template <bool XH>
__device__ __forceinline__ void sparse3(const P2 &p, const uint32_t base, const int spw, const int lane, const uint32_t boxtag) {
    const int N = p.N;
    const uint32_t xs_u32 = base + OFF_X;
    // ---------------- static: dense-row slice ----------------
    float hfr[HYB_R2];
    const bool hyb_on = p.full_rows && spw == 0 && (int)blockIdx.x < p.hc;
    const bool hyb_multi = hyb_on && p.topX <= 32;
    int kb = 0, ke = 0, nsl = 1, rs = 0, hj = lane;
    if (hyb_on) { kb = blockIdx.x * p.hrows; ke = min(p.K, kb + p.hrows); }
    if (hyb_multi) {
        nsl = 32 / p.topX;
        rs = lane / p.topX;
        hj = lane - rs * p.topX;
#pragma unroll
        for (int i = 0; i < HYB_R2; ++i) {
            const int k = kb + rs + nsl * i;
            hfr[i] = (rs < nsl && k < ke) ? __ldg(p.full_rows + (size_t)k * p.topX + hj) : 0.f;
        }
    }
    // ---------------- static: this warp's CSR rows [r, rb) of the CTA's [ca, cb) ----------------
    int r = 0, rb = 0;
    if (p.rows) {
        const int ca = min(N, (int)blockIdx.x * p.csr_rpc), cb = min(N, ca + p.csr_rpc);
        // warp 0 has the dense rows (atomics + a fence before its announcement: ~2-3 us) and then takes no CSR rows: outlier sums are what
        // the owners of ALL strips wait for, they must not queue behind that fence
        const int tot = cb - ca, w0 = p.full_rows ? 0 : tot / NSPW, rest = tot - w0;
        r = spw == 0 ? ca : ca + w0 + (int)((long long)rest * (spw - 1) / (NSPW - 1));
        rb = spw == 0 ? ca + w0 : ca + w0 + (int)((long long)rest * spw / (NSPW - 1));
    }
    const int nr = rb - r;
    int LPR = 4;
    if (nr > 0) {
        const int e_lo = __ldg(p.rows + r), e_hi = __ldg(p.rows + rb);
        const int avg = (e_hi - e_lo) / nr;
        LPR = avg > 128 ? 32 : avg > 64 ? 16 : avg > 32 ? 8 : avg > 8 ? 4 : 2;
    }
    const int RPP = 32 / LPR, part = lane & (LPR - 1), rslot = lane / LPR;
    const int npass = (nr + RPP - 1) / RPP;
    constexpr int T = 8;  // (col, val) pairs per lane and pass
    auto load_ptrs = [&](int pass, int &a0, int &a1) {
        const int row = r + pass * RPP + rslot;
        a0 = a1 = 0;
        if (pass < npass && row < rb) { a0 = __ldg(p.rows + row); a1 = __ldg(p.rows + row + 1); }
    };
    auto load_elems = [&](int a0, int a1, int (&cc)[T], float (&vv)[T]) {
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int e = a0 + part + t * LPR;
            const bool ok = e < a1;
            cc[t] = ok ? __ldg(p.cols + e) : 0;
            vv[t] = ok ? __ldg(p.vals + e) : 0.f;
        }
    };
    int a0c, a1c, a0n, a1n, a0nn, a1nn;
    int cc[T], cn[T];
    float vv[T], vn[T];
    load_ptrs(0, a0c, a1c);
    load_ptrs(1, a0n, a1n);
    load_elems(a0c, a1c, cc, vv);

    named_bar_sync(2, NCT + NSPW * 32);  // x is in shared memory

    // ---------------- dense rows (v2's phase B) ----------------
    if (hyb_on) {
        for (int jb = 0; jb < (hyb_multi ? 1 : p.topX); jb += 32) {
            float a = 0.f;
            int j;
            if (hyb_multi) {
                j = hj;
#pragma unroll
                for (int i = 0; i < HYB_R2; ++i) {
                    const int k = kb + rs + nsl * i;
                    if (rs < nsl && k < ke) a += hfr[i] * xs_load<XH>(xs_u32, k);
                }
                for (int k = kb + rs + nsl * HYB_R2; rs < nsl && k < ke; k += nsl)
                    a += __ldg(p.full_rows + (size_t)k * p.topX + hj) * xs_load<XH>(xs_u32, k);
                for (int sl = 1; sl < nsl; ++sl) {
                    const float v = __shfl_sync(0xffffffffu, a, (hj + sl * p.topX) & 31);
                    if (rs == 0) a += v;
                }
                if (rs != 0) j = p.topX;
                // dense rows that feed the SAME channel travel as one word (the first of them carries the sum): the owner of that channel's
                // strip then reads hc words, not hc per row - a checkpoint without dense rows loads as topX rows on channel 0 (llama.py:182)
                {
                    const int cme = (rs == 0 && hj < p.topX) ? __ldg(p.fri + hj) : -1 - lane;
                    bool first = true;
                    float tot = a;
                    for (int j2 = 0; j2 < p.topX; ++j2) {
                        const int c2 = __shfl_sync(0xffffffffu, cme, j2);
                        const float a2 = __shfl_sync(0xffffffffu, a, j2);
                        if (c2 == cme && j2 < hj) first = false;
                        if (c2 == cme && j2 > hj) tot += a2;
                    }
                    a = tot;
                    if (!first) j = p.topX;
                }
            } else {
                j = jb + lane;
                if (j < p.topX) {
                    const float *fr = p.full_rows + (size_t)kb * p.topX + j;
                    for (int k = kb; k < ke; ++k, fr += p.topX) a += __ldg(fr) * xs_load<XH>(xs_u32, k);
                }
            }
            // this CTA's part of dense row j: one tagged word, no fence, no flag (the owner of the channel's strip sums the hc words of row j)
            if (j < p.topX) st_relaxed_u64(p.ws_dbox + (size_t)j * MAX_GRID_V2 + blockIdx.x, box_word(a, boxtag));
        }
    }
    // ---------------- CSR passes ----------------
    for (int pass = 0; pass < npass; ++pass) {
        load_ptrs(pass + 2, a0nn, a1nn);
        load_elems(a0n, a1n, cn, vn);
        float s0 = 0.f, s1 = 0.f;
        {
            float xv[T];
#pragma unroll
            for (int t = 0; t < T; ++t) xv[t] = xs_load<XH>(xs_u32, cc[t]);   // (padding slots read x[0] with val 0)
#pragma unroll
            for (int t = 0; t < T; t += 2) { s0 = fmaf(vv[t], xv[t], s0); s1 = fmaf(vv[t + 1], xv[t + 1], s1); }
        }
        float tot = s0 + s1;
        for (int d = 1; d < LPR; d <<= 1) tot += __shfl_xor_sync(0xffffffffu, tot, d);
        // rows longer than T * LPR: the whole warp sums the remainder, row after row
        unsigned longm = __ballot_sync(0xffffffffu, part == 0 && a1c - a0c > T * LPR);
        while (longm) {
            const int i = __ffs(longm) - 1;
            longm &= longm - 1;
            const int b0 = __shfl_sync(0xffffffffu, a0c, i) + T * LPR, b1 = __shfl_sync(0xffffffffu, a1c, i);
            float a = 0.f, b = 0.f;
            for (int e = b0 + lane; e < b1; e += 128) {
                const int c0 = __ldg(p.cols + e), c1 = e + 32 < b1 ? __ldg(p.cols + e + 32) : 0, c2 = e + 64 < b1 ? __ldg(p.cols + e + 64) : 0,
                          c3 = e + 96 < b1 ? __ldg(p.cols + e + 96) : 0;
                const float v0 = __ldg(p.vals + e), v1 = e + 32 < b1 ? __ldg(p.vals + e + 32) : 0.f, v2 = e + 64 < b1 ? __ldg(p.vals + e + 64) : 0.f,
                            v3 = e + 96 < b1 ? __ldg(p.vals + e + 96) : 0.f;
                a = fmaf(v0, xs_load<XH>(xs_u32, c0), a); b = fmaf(v1, xs_load<XH>(xs_u32, c1), b);
                a = fmaf(v2, xs_load<XH>(xs_u32, c2), a); b = fmaf(v3, xs_load<XH>(xs_u32, c3), b);
            }
            a = warp_sum(a + b);
            if (lane == i) tot += a;
        }
        const int row = r + pass * RPP + rslot;
        if (part == 0 && row < rb) st_relaxed_u64(p.ws_cbox + row, box_word(tot, boxtag));
        a0c = a0n; a1c = a1n; a0n = a0nn; a1n = a1nn;
#pragma unroll
        for (int t = 0; t < T; ++t) { cc[t] = cn[t]; vv[t] = vn[t]; }
    }
}

template <int BITS, int MODE, bool MULTI>
__global__ void __launch_bounds__(THREADS2, 1) lutgemv_seq_kernel(const SeqCfg c) {
    using C = C2<BITS, MODE>;
    constexpr bool XH = true;
    constexpr int NTB = C::TAB <= 16384 ? 4 : 2;  // table buffers (and strip accumulators): 4 x 4 KB exact, 2 x 64 KB for the 4-bit pair table
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    if (raw != c.smem_raw) __trap();
    uint32_t base = (raw + 127u) & ~127u;
    asm volatile("mov.u32 %0, %0;" : "+r"(base));
    unsigned char *sm = smem_raw + (base - raw);
    const uint32_t bar_u32 = base + OFF_BAR;   // full[s] +8s, empty[s] +128+8s, tfull[b] +256+8b, tfree[b] +288+8b
    const uint32_t xs_u32 = base + OFF_X;
    const uint32_t sacc0 = xs_u32 + (uint32_t)c.xbytes;                 // float [NTB][64]
    const uint32_t lo_base = sacc0 + NTB * STRIP * 4;
    const uint32_t tab0 = (lo_base + (uint32_t)C::TAB - 1u) & ~((uint32_t)C::TAB - 1u);
    const int nstage = c.nstage;
    const int n_lo = min(nstage, (int)((tab0 - lo_base) / C::STAGE));
    const uint32_t hi_base = tab0 + NTB * C::TAB - (uint32_t)n_lo * C::STAGE;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(bar_u32 + 8 * s, 1);
            mbar_init(bar_u32 + 128 + 8 * s, NWC);
        }
        for (int b = 0; b < NTB; ++b) {
            mbar_init(bar_u32 + 256 + 8 * b, NBW);
            mbar_init(bar_u32 + 288 + 8 * b, NWC);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t epoch = *reinterpret_cast<const volatile unsigned *>(c.epoch);
    const uint32_t xtag = (epoch & 0x7fffu) | 0x8000u;   // 16-bit tag of this token's activation words
    const uint32_t xwant = xtag << 16;
    const int ngemv = c.ngemv;

    // the CTA's part of GEMV `p`: units [g0, g0 + len) of the strip-major [strip][unit] space
#define SEQ_GEOMETRY(p)                                                                   \
    const int N = (p).N, R = (p).R;                                                       \
    const int g0 = min((int)blockIdx.x * (p).chunk, (p).T), g1 = min(g0 + (p).chunk, (p).T); \
    const int len = g1 - g0;                                                              \
    const int s0 = g0 / R, r0 = g0 - s0 * R;                                              \
    const int nseg = len > 0 ? (g1 - 1) / R - s0 + 1 : 0;                                 \
    (void)N; (void)r0; (void)s0; (void)nseg;

    if (warp == WARP_PROD) {
        // =========================== TMA producer: runs ahead of the consumers across GEMV boundaries ===========================
        const uint64_t pol = l2_evict_first_policy();
        int slot = 0;
        uint32_t ph = 0;
        bool refill = false;
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&c.descs[0].tm_big)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&c.descs[0].tm_small)) : "memory");
        }
        for (int g = 0; g < ngemv; ++g) {
            const SeqDesc &d = c.descs[g];
            const P2 &p = d.p;
            SEQ_GEOMETRY(p)
            if (lane == 0 && g + 1 < ngemv) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&c.descs[g + 1].tm_big)) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&c.descs[g + 1].tm_small)) : "memory");
            }
            {   // a 1/grid slice of the GEMV's look-up table and row pointers -> L2, for everybody (cold otherwise: a token never finds them there)
                const int G = (int)gridDim.x;
                const size_t lut_lines = ((size_t)N * C::L * 4 + 127) / 128, per = (lut_lines + G - 1) / G;
                for (size_t i = (size_t)blockIdx.x * per + lane; i < min(lut_lines, ((size_t)blockIdx.x + 1) * per); i += 32)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(p.lut) + 128 * i));
                if (p.rows) {
                    const size_t row_lines = ((size_t)(N + 1) * 4 + 127) / 128, rper = (row_lines + G - 1) / G;
                    for (size_t i = (size_t)blockIdx.x * rper + lane; i < min(row_lines, ((size_t)blockIdx.x + 1) * rper); i += 32)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(p.rows) + 128 * i));
                }
            }
            for (int seg = 0; seg < nseg; ++seg) {
                const int sstart = seg == 0 ? 0 : seg * R - r0, send = min(len, (seg + 1) * R - r0);
                const int seglen = send - sstart;
                const int u0 = r0 + sstart - seg * R;
                const int col0 = (s0 + seg) * STRIP;
                for (int u = 0; u < seglen; u += SU2) {
                    if (refill) mbar_wait(bar_u32 + 128 + 8 * slot, ph);
                    const int nu = min(SU2, seglen - u);
                    const uint32_t full = bar_u32 + 8 * slot;
                    const uint32_t dst = (slot < n_lo ? lo_base : hi_base) + slot * C::STAGE;
                    const int row0 = (u0 + u) * C::ROWS;
                    if (lane == 0) mbar_expect_tx(full, (uint32_t)nu * C::UNIT);
                    __syncwarp();
                    if (nu == SU2) {
                        if (lane == 0) tma_tile2d_g2s(dst, &d.tm_big, col0, row0, full, pol);
                    } else if (2 * lane < nu) {
                        tma_tile2d_g2s(dst + lane * 2 * C::UNIT, &d.tm_small, col0, row0 + lane * 2 * C::ROWS, full, pol);
                    }
                    if (++slot == nstage) {
                        slot = 0;
                        if (refill) ph ^= 1u;
                        refill = true;
                    }
                }
            }
        }
    } else if (warp >= WARP_SP) {
        // =========================== outlier warps (v2::sparse2, once per GEMV) ===========================
        for (int g = 0; g < ngemv; ++g) {
            const P2 &p = c.descs[g].p;
            sparse3<XH>(p, base, warp - WARP_SP, lane, (epoch << 10) | (uint32_t)(g + 1));
            TRACE(10, lane == 0);
            if (g + 1 < ngemv) bar_arrive(6, NCT + NSPW * 32);  // done reading this GEMV's x from shared memory
        }
    } else if (warp >= WARP_BLD) {
        // =========================== table builders + strip finishers (2 warps, thread = column slot) ===========================
        const int bt = tid - WARP_BLD * 32;
        const uint32_t lutbuf = base + OFF_LUT;
#pragma unroll
        for (int b = 0; b < NTB; ++b) sts_u32(sacc0 + 4 * (b * STRIP + bt), 0u);
        int *const err = c.err;
        int q0 = 0;  // segments of earlier GEMVs: segment s of this one uses table buffer / accumulator (q0 + s) % NTB
        {
            const P2 &p = c.descs[0].p;
            SEQ_GEOMETRY(p)
            if (nseg > 0) lut_prefetch<BITS>(p, lutbuf, s0, bt);
            if (nseg > 1) lut_prefetch<BITS>(p, lutbuf + LUTBUF, s0 + 1, bt);
        }
        for (int g = 0; g < ngemv; ++g) {
            const SeqDesc &d = c.descs[g];
            const P2 &p = d.p;
            SEQ_GEOMETRY(p)
            const uint32_t boxtag = (epoch << 10) | (uint32_t)(g + 1);
            float *const acc_out = p.ws_acc;
            int *const flags = p.ws_cnt + 64;
            auto sacc_of = [&](int s) { return sacc0 + 4u * (uint32_t)(((q0 + s) % NTB) * STRIP + bt); };
            auto tfree_wait = [&](int s) { mbar_wait(bar_u32 + 288 + 8 * ((q0 + s) % NTB), (uint32_t)(((q0 + s) / NTB) & 1)); };
            // Strip sums of segment s go to the strip's owner: another CTA (mailbox row) if the strip started before our range, else ourselves -
            // and then they never leave the SM: sown[u][column] for the u-th owned strip (a strip inside one CTA is exactly one segment).
            // (v2 and the first sequence build sent them through the global accumulator - a RED and, in the finishing pass, a load that has to
            // wait for it: 1.5 us of L2 round trips on the critical path between the last weight and y, profiles/r02_seq_trace_2.txt.)
            const uint32_t sown = base + OFF_SROW;  // float [SOWN][64] (the row arrays of v2's outlier warps are not used by the sequence kernel)
            constexpr int SOWN = (2 * (SP_ROWS + 1) * 4) / (STRIP * 4);
            const int ofs = r0 != 0 ? 1 : 0;        // segment s is owned strip s - ofs
            auto flush = [&](int s) {
                const uint32_t a = sacc_of(s);
                const float v = lds_f32(a);
                sts_u32(a, 0u);
                const int col = (s0 + s) * STRIP + bt;
                if (s == 0 && r0 != 0) st_relaxed_u64(p.ws_hbox + (size_t)blockIdx.x * STRIP + bt, box_word(v, boxtag));
                else if (s - ofs < SOWN) sts_u32(sown + 4u * (uint32_t)((s - ofs) * STRIP + bt), __float_as_uint(v));
                else if (col < N) atomicAdd(acc_out + col, v);
            };
            for (int s = 0; s < nseg; ++s) {
                const int b = (q0 + s) % NTB;
                if (s >= NTB) {
                    tfree_wait(s - NTB);
                    flush(s - NTB);
                }
                if (s + 1 < nseg) cp_async_wait_pending<1>();
                else cp_async_wait_all();
                named_bar_sync(3, NBT);
                build_table<BITS, MODE>(tab0 + b * C::TAB, lutbuf + (s & 1) * LUTBUF, bt);
                named_bar_sync(3, NBT);
                if (lane == 0) mbar_arrive(bar_u32 + 256 + 8 * b);
                if (s + 2 < nseg) lut_prefetch<BITS>(p, lutbuf + (s & 1) * LUTBUF, s0 + s + 2, bt);
            }
            if (g + 1 < ngemv) {  // the next GEMV's first raw LUT rows travel while this one is being finished
                const P2 &pn = c.descs[g + 1].p;
                const int gn0 = min((int)blockIdx.x * pn.chunk, pn.T), gn1 = min(gn0 + pn.chunk, pn.T);
                const int sn0 = gn0 / pn.R, nsegn = gn1 > gn0 ? (gn1 - 1) / pn.R - sn0 + 1 : 0;
                if (nsegn > 0) lut_prefetch<BITS>(pn, lutbuf, sn0, bt);
                if (nsegn > 1) lut_prefetch<BITS>(pn, lutbuf + LUTBUF, sn0 + 1, bt);
            }
            // ---- finish the strips this CTA owns (those that start in its range); see lutgemv_v2.cuh for the protocol ----
            TRACE(12, bt == 0);
            const long long cb0 = (long long)blockIdx.x * p.chunk, cb1 = min((long long)p.T, cb0 + p.chunk);
            const int so0 = (int)((cb0 + R - 1) / R), so1 = (int)((cb1 + R - 1) / R), nown = so1 - so0;
            const bool last_owned = nseg >= 2 || r0 == 0;
            constexpr int FAST = 4;
            const int lidx = last_owned && nseg > 0 ? nown - 1 : -1;
            const int nh = lidx >= 0 ? (int)((((long long)(so1 - 1) + 1) * R - 1) / p.chunk) - (int)blockIdx.x : 0;
            // Dense rows whose channel lies in a strip we own: listed now (static), summed below.  sdj[i] = dense row, sden[i] = its sum.
            int *const sdj = reinterpret_cast<int *>(sm + OFF_CSR);
            const uint32_t sden = base + OFF_CSR + 4 * 128, sdn_a = base + OFF_CSR + 8 * 128;
            if (p.full_rows) {
                // (one warp, registers and shuffles: a serial scan of full_row_indices by one thread cost 435 dependent loads = 6 us for the 30
                //  dense rows of a stacked q/k/v layer - on every CTA, before its first mailbox row went out)
                const int topX = p.topX;
                if (bt == 0) sts_u32(sdn_a, 0u);
                sts_u32(sden + 4 * bt, 0u);
                sts_u32(sden + 4 * (bt + NBT), 0u);
                named_bar_sync(3, NBT);
                if (topX <= 32) {
                    if (bt < 32) {
                        const int cc = bt < topX ? __ldg(p.fri + bt) : -1 - bt;
                        bool first = true;  // rows on the same channel were combined by the contributors: the first of them carries the sum
                        for (int j2 = 0; j2 < topX; ++j2) {
                            const int c2 = __shfl_sync(0xffffffffu, cc, j2);
                            if (c2 == cc && j2 < bt) first = false;
                        }
                        const bool mine = bt < topX && first && cc >= 0 && cc < N && cc / STRIP >= so0 && cc / STRIP < so1;
                        const unsigned m = __ballot_sync(0xffffffffu, mine);
                        if (mine) sdj[__popc(m & ((1u << bt) - 1u))] = bt;
                        if (bt == 0) sts_u32(sdn_a, (uint32_t)__popc(m));
                    }
                } else {
                    for (int j = bt; j < topX; j += NBT) {
                        const int cc = __ldg(p.fri + j);
                        if (cc >= 0 && cc < N && cc / STRIP >= so0 && cc / STRIP < so1) {
                            uint32_t at;
                            asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(at) : "r"(sdn_a) : "memory");
                            sdj[at] = j;
                        }
                    }
                }
                named_bar_sync(3, NBT);
            }
            const int sdn = p.full_rows ? (int)lds_u32(sdn_a) : 0;
            for (int s = max(0, nseg - NTB); s < nseg - 1; ++s) {
                tfree_wait(s);
                flush(s);
            }
            TRACE(13, bt == 0);
            {
                unsigned long long cw[FAST], hw[3];
#pragma unroll
                for (int u = 0; u < FAST; ++u) {
                    const int col = (so0 + u) * STRIP + bt;
                    cw[u] = (p.rows && u < nown && col < N) ? ld_relaxed_u64(p.ws_cbox + col) : box_word(0.f, boxtag);
                }
#pragma unroll
                for (int h = 0; h < 3; ++h)
                    hw[h] = h < nh ? ld_relaxed_u64(p.ws_hbox + ((size_t)blockIdx.x + 1 + h) * STRIP + bt) : box_word(0.f, boxtag);
                // dense rows: hc tagged words per listed row (one from every contributing CTA), four loads in flight per thread; row sums are
                // collected in shared memory and added to the column they belong to by the thread that owns that column
                float last_dense = 0.f;
                if (sdn > 0) {
                    const int hc = p.hc;
                    unsigned long long *const dbox = p.ws_dbox;
                    for (int j0 = 0; j0 < sdn; j0 += 4) {  // four rows x up to four words per thread in flight (hc <= MAX_GRID_V2 = 4 x NBT)
                        unsigned long long dw[4][4];
                        int row[4];
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            row[a] = j0 + a < sdn ? sdj[j0 + a] : -1;
#pragma unroll
                            for (int b = 0; b < 4; ++b)
                                if (row[a] >= 0 && bt + b * NBT < hc) dw[a][b] = ld_relaxed_u64(dbox + (size_t)row[a] * MAX_GRID_V2 + bt + b * NBT);
                        }
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            if (row[a] < 0) continue;  // (uniform)
                            float v = 0.f;
#pragma unroll
                            for (int b = 0; b < 4; ++b)
                                if (bt + b * NBT < hc) v += box_take(dw[a][b], dbox + (size_t)row[a] * MAX_GRID_V2 + bt + b * NBT, err, boxtag, true);
                            v = warp_sum(v);
                            if (lane == 0) asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(sden + 4u * (uint32_t)(j0 + a)), "f"(v) : "memory");
                        }
                    }
                    named_bar_sync(3, NBT);
                    for (int i = 0; i < sdn; ++i) {
                        const int cc = __ldg(p.fri + sdj[i]);
                        if ((cc & (STRIP - 1)) == bt) {
                            const int u = cc / STRIP - so0;
                            const float v = lds_f32(sden + 4u * (uint32_t)i);
                            if (u == lidx) last_dense += v;
                            else if (u < SOWN) {
                                const uint32_t a = sown + 4u * (uint32_t)(u * STRIP + bt);
                                sts_u32(a, __float_as_uint(lds_f32(a) + v));
                            } else atomicAdd(acc_out + cc, v);
                        }
                    }
                    named_bar_sync(3, NBT);  // (the list and the sums are reused by the next item; an idle CTA has no other barrier before that)
                }
                const int w = MULTI ? N / p.xw_members : 0;
                auto store_y = [&](int col, float yv) {
                    if (p.bias) yv += __ldg(p.bias + col);
                    const uint32_t word = (xtag << 16) | (uint32_t)__half_as_ushort(__float2half_rn(yv));
                    if constexpr (!MULTI) {
                        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(c.arena_base + d.y_off + 4ull * (unsigned)col), "r"(word) : "memory");
                    } else {
                        // local column col of the stacked shard = column j of member m -> element [m][rank*w + j] of the full-length vector, in EVERY rank's arena
                        const int mm = col / w, j = col - mm * w;
                        const unsigned long long e = (unsigned long long)mm * p.xw_nfull + (unsigned long long)c.rank * w + j;
                        for (int pr = 0; pr < c.world; ++pr)
                            asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(__ldg(c.peer_base + pr) + d.y_off + 4ull * e), "r"(word) : "memory");
                    }
                    if (p.out) reinterpret_cast<__half *>(p.out)[col] = __float2half_rn(yv);
                };
                // per owned strip but the last segment's: [our own sums (+ dense rows), from shared memory] + [the column's outlier sum]
                TRACE(14, bt == 0);
                float last_others = last_dense;
#pragma unroll
                for (int u = 0; u < FAST; ++u) {
                    const int col = (so0 + u) * STRIP + bt;
                    if (u < nown && col < N) {
                        float yv = 0.f;
                        if (u != lidx && u < SOWN) yv = lds_f32(sown + 4u * (uint32_t)(u * STRIP + bt));
                        if (p.rows) yv += box_take(cw[u], p.ws_cbox + col, err, boxtag, true);
                        if (u == lidx) last_others += yv;
                        else store_y(col, yv);
                    }
                }
                for (int i = FAST; i < nown; ++i) {  // (more owned strips than FAST; beyond SOWN the sums went through the global accumulator)
                    const int col = (so0 + i) * STRIP + bt;
                    if (col < N) {
                        float yv = 0.f;
                        if (i >= SOWN) { yv = __ldcg(p.ws_acc + col); p.ws_acc[col] = 0.f; }
                        else if (i != lidx) yv = lds_f32(sown + 4u * (uint32_t)(i * STRIP + bt));
                        if (p.rows) yv += box_take(0ull, p.ws_cbox + col, err, boxtag, true);
                        if (i == lidx) last_others += yv;
                        else store_y(col, yv);
                    }
                }
                TRACE(15, bt == 0);
#pragma unroll
                for (int h = 0; h < 3; ++h)
                    if (h < nh) last_others += box_take(hw[h], p.ws_hbox + ((size_t)blockIdx.x + 1 + h) * STRIP + bt, err, boxtag, true);
                // (a strip that spans many CTAs - narrow column shards: 18 CTAs per strip for a 512-column shard of down_proj - has many such
                //  words: eight reads in flight, not one round trip after the other)
                for (int h0 = 3; h0 < nh; h0 += 8) {
                    unsigned long long mw[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (h0 + k < nh) mw[k] = ld_relaxed_u64(p.ws_hbox + ((size_t)blockIdx.x + 1 + h0 + k) * STRIP + bt);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (h0 + k < nh) last_others += box_take(mw[k], p.ws_hbox + ((size_t)blockIdx.x + 1 + h0 + k) * STRIP + bt, err, boxtag, true);
                }
                TRACE(8, bt == 0);
                if (nseg > 0) {
                    const int s = nseg - 1;
                    tfree_wait(s);
                    if (last_owned) {
                        const uint32_t a = sacc_of(s);
                        const float own = lds_f32(a);
                        sts_u32(a, 0u);
                        const int lcol = (s0 + s) * STRIP + bt;
                        if (lcol < N) store_y(lcol, last_others + own);
                    } else {
                        flush(s);
                    }
                }
            }
            TRACE(9, bt == 0);
            q0 += nseg;
        }
    } else {
        // =========================== consumers ===========================
        const int i16 = lane & 15, jsel = lane >> 4;
        Acc A;
#pragma unroll
        for (int t = 0; t < 4; ++t) A.a[t] = 0ull;
#pragma unroll
        for (int t = 0; t < 8; ++t) A.f[t] = 0.f;
        const uint32_t slotb[4] = {(uint32_t)((((0 ^ jsel) << 4) | i16) << 2), (uint32_t)((((1 ^ jsel) << 4) | i16) << 2),
                                   (uint32_t)((((2 ^ jsel) << 4) | i16) << 2), (uint32_t)((((3 ^ jsel) << 4) | i16) << 2)};
        const uint32_t lane_in_stage = (uint32_t)((2 * warp + jsel) * C::UNIT + i16 * 16);
        constexpr int XB = C::XU * 2;
        constexpr int XBS = SU2 * XB;
        int slot = 0, q0 = 0;
        uint32_t par = 0;
        auto stage_of = [&](int sl) { return (sl < n_lo ? lo_base : hi_base) + sl * C::STAGE + lane_in_stage; };
        auto advance = [&]() { if (++slot == nstage) { slot = 0; par ^= 1u; } };
        for (int g = 0; g < ngemv; ++g) {
            const SeqDesc &d = c.descs[g];
            const P2 &p = d.p;
            SEQ_GEOMETRY(p)
            TRACE(2, tid == 0);
            if (g > 0) named_bar_sync(6, NCT + NSPW * 32);  // every consumer warp and the outlier warps are done with the previous x
            if (d.x_tag == nullptr) {
                const int n16 = p.K * 2 / 16;
                for (int e = tid; e < n16; e += NCT) cp_async16(xs_u32 + 16 * e, reinterpret_cast<const unsigned char *>(p.x) + 16 * (size_t)e);
                cp_async_commit();
                cp_async_wait_all();
            } else {
                poll_tagged<MULTI>(d.x_tag, p.K / 4, tid, NCT, xwant, c.err, [&](int i, uint32_t lo, uint32_t hi) {
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(xs_u32 + 8u * (uint32_t)i), "r"(lo), "r"(hi) : "memory");
                });
            }
            named_bar_sync(2, NCT + NSPW * 32);  // x visible to the consumers and the outlier warps
            TRACE(4, tid == 0);
            for (int seg = 0; seg < nseg; ++seg) {
                const int q = q0 + seg, b = q % NTB;
                const uint32_t tab = tab0 + b * C::TAB;
                const uint32_t tb_hi = (MODE == 1 && BITS == 4) ? tab : (BITS == 4 ? (tab & 0xFFFF0000u) : tab);
                uint32_t l[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    l[t] = tb_hi | slotb[t];
                    asm volatile("mov.u32 %0, %0;" : "+r"(l[t]));
                }
                const uint32_t segc = ((tab >> 8) & 0xF0u) * 0x01010101u;
                const int sstart = seg == 0 ? 0 : seg * R - r0, send = min(len, (seg + 1) * R - r0);
                const int seglen = send - sstart;
                const int u0 = r0 + sstart - seg * R;
                uint32_t xaddr = xs_u32 + (uint32_t)((u0 + 2 * warp + jsel) * XB);
                const int nstg = (seglen + SU2 - 1) / SU2;
                const int mine = 2 * warp < seglen ? (seglen - 2 * warp + SU2 - 1) / SU2 : 0;
                mbar_wait(bar_u32 + 256 + 8 * b, (uint32_t)((q / NTB) & 1));
                TRACE(16 + (seg < 7 ? seg : 7), tid == 0);
#if SQLLM_SEQ_PREFETCH
                // The words (and x slice) of stage k+1 are requested BEFORE the gathers of stage k are issued: one of the two serialized
                // shared-memory round trips per position (word fetch, then gathers - each queues behind the other warps' gathers in the LSU)
                // disappears from a warp's critical path.  One copy of the math, the look-ahead lives in a second register set that is
                // moved over at the end of the trip (8 / 12 MOVs per ~140 instructions); the loop is kept rolled (i-cache: see v2).
                {
                    Fetch<BITS, XH> F, Fn;
                    mbar_wait(bar_u32 + 8 * slot, par);
                    if (0 < mine) fetch2<BITS, XH>(F, stage_of(slot), xaddr);
#pragma unroll 1
                    for (int k = 0; k < nstg; ++k) {
                        const int cur = slot;
                        advance();
                        if (k + 1 < nstg) {
                            mbar_wait(bar_u32 + 8 * slot, par);
                            if (k + 1 < mine) fetch2<BITS, XH>(Fn, stage_of(slot), xaddr + XBS);
                        }
                        if (k < mine) math2<BITS, MODE, XH>(F, jsel, l, segc, xaddr, A);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_u32 + 128 + 8 * cur);
                        F = Fn;
                        xaddr += XBS;
                    }
                }
#else
                for (int k = 0; k < nstg; ++k) {
                    mbar_wait(bar_u32 + 8 * slot, par);
                    if (k < mine) {
                        Fetch<BITS, XH> F;
                        fetch2<BITS, XH>(F, stage_of(slot), xaddr);
                        math2<BITS, MODE, XH>(F, jsel, l, segc, xaddr, A);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_u32 + 128 + 8 * slot);
                    advance();
                    xaddr += XBS;
                }
#endif
                {
                    float s[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        s[t] = MODE == 0 ? sum2(A.a[t]) : A.f[t] + A.f[t + 4];
                        A.a[t] = 0ull;
                        A.f[t] = 0.f;
                        A.f[t + 4] = 0.f;
                    }
                    const float v0 = __shfl_xor_sync(0xffffffffu, s[1], 16);
                    const float v1 = __shfl_xor_sync(0xffffffffu, s[0], 16);
                    const float v2 = __shfl_xor_sync(0xffffffffu, s[3], 16);
                    const float v3 = __shfl_xor_sync(0xffffffffu, s[2], 16);
                    if (jsel == 0 && mine > 0) {
                        const uint32_t a = sacc0 + (uint32_t)(b * STRIP + 4 * i16) * 4u;
                        asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(s[0] + v0) : "memory");
                        asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a + 4), "f"(s[1] + v1) : "memory");
                        asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a + 8), "f"(s[2] + v2) : "memory");
                        asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a + 12), "f"(s[3] + v3) : "memory");
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_u32 + 288 + 8 * b);
                }
                TRACE(24 + (seg < 7 ? seg : 7), tid == 0);
            }
            TRACE(6, tid == 0);
            q0 += nseg;
        }
        // exports: plain fp16 copies of the vectors the caller asked for (on several GPUs this is also where the last exchange completes)
        for (int e = (int)blockIdx.x; e < c.nexport; e += (int)gridDim.x) {
            const SeqExport ex = c.exports[e];
            poll_tagged<MULTI>(ex.src, ex.n / 4, tid, NCT, xwant, c.err, [&](int i, uint32_t lo, uint32_t hi) {
                *reinterpret_cast<uint2 *>(reinterpret_cast<unsigned char *>(ex.dst) + 8 * (size_t)i) = make_uint2(lo, hi);
            });
        }
    }
#undef SEQ_GEOMETRY
}

}  // namespace seq
