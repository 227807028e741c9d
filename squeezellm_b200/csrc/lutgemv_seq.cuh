// lutgemv_seq.cuh - the SEQUENCE kernel: one persistent launch runs a whole list of dependent LUT-GEMVs (a decode token's QuantLinear
// matvecs: llama.py:226-234 calls them one by one, 224 launches per LLaMA-7B token).  Included by lutgemv_kernels.cu after lutgemv_v2.cuh;
// the per-position math, table builders, outlier warps and mailboxes are v2's (namespace v2), what is new is everything BETWEEN two GEMVs.
//
// Why: with one launch per GEMV (v2 + PDL + CUDA graph) a 4096x4096 layer took 8.9 us in the chain for 1.5 us of HBM time; a fit over the
// four 7B shapes gave ~6 us of fixed cost per launch (CTA exit -> grid completion -> dependency release -> CTA start, barrier init, tensor
// map fetch, x staging, ring fill from cold) - 128 launches per token, 40 % of the step (profiles/r02_bench_default.json).  Here
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
            named_bar_sync(5, NSPW * 32);  // the CTA-wide row arrays of the previous GEMV have been published by all four warps
            sparse2<XH, true>(p, sm, base, warp - WARP_SP, lane, p.ws_acc, (epoch << 10) | (uint32_t)(g + 1));
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
            auto flush = [&](int s) {  // strip sums of segment s: to the strip's owner (mailbox row) or into our own accumulator; sums back to zero
                const uint32_t a = sacc_of(s);
                const float v = lds_f32(a);
                sts_u32(a, 0u);
                const int col = (s0 + s) * STRIP + bt;
                if (s == 0 && r0 != 0) st_relaxed_u64(p.ws_hbox + (size_t)blockIdx.x * STRIP + bt, box_word(v, boxtag));
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
            for (int s = max(0, nseg - NTB); s < nseg - 1; ++s) {
                tfree_wait(s);
                flush(s);
            }
            const long long cb0 = (long long)blockIdx.x * p.chunk, cb1 = min((long long)p.T, cb0 + p.chunk);
            const int so0 = (int)((cb0 + R - 1) / R), so1 = (int)((cb1 + R - 1) / R), nown = so1 - so0;
            const bool last_owned = nseg >= 2 || r0 == 0;
            {
                constexpr int FAST = 4;
                const int lidx = last_owned && nseg > 0 ? nown - 1 : -1;
                const int nh = lidx >= 0 ? (int)((((long long)(so1 - 1) + 1) * R - 1) / p.chunk) - (int)blockIdx.x : 0;
                unsigned long long cw[FAST], hw[3];
#pragma unroll
                for (int u = 0; u < FAST; ++u) {
                    const int col = (so0 + u) * STRIP + bt;
                    cw[u] = (p.rows && u < nown && col < N) ? ld_relaxed_u64(p.ws_cbox + col) : box_word(0.f, boxtag);
                }
#pragma unroll
                for (int h = 0; h < 3; ++h)
                    hw[h] = h < nh ? ld_relaxed_u64(p.ws_hbox + ((size_t)blockIdx.x + 1 + h) * STRIP + bt) : box_word(0.f, boxtag);
                if (p.full_rows) {
                    for (int i = bt; i < nown; i += NBT) {
                        const int strip = so0 + i;
                        bool hch = false;
                        for (int j = 0; j < p.topX; ++j) {
                            const int cc = __ldg(p.fri + j);
                            hch |= (cc >= 0 && cc < N && cc / STRIP == strip);
                        }
                        if (hch) {
                            int seen;
                            unsigned long long t0 = 0ull, t1;
                            do {
                                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(flags + strip) : "memory");
                                if (seen >= p.hc) break;
                                __nanosleep(40);
                                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                                if (t0 == 0ull) t0 = t1;
                                if (t1 - t0 > 2000000000ull) { *reinterpret_cast<volatile int *>(err) = 1; break; }
                            } while (true);
                            flags[strip] = 0;
                        }
                    }
                    named_bar_sync(3, NBT);
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
                float av[FAST];
#pragma unroll
                for (int u = 0; u < FAST; ++u) {
                    const int col = (so0 + u) * STRIP + bt;
                    av[u] = (u < nown && col < N) ? __ldcg(p.ws_acc + col) : 0.f;
                }
                float last_others = 0.f;
#pragma unroll
                for (int u = 0; u < FAST; ++u) {
                    const int col = (so0 + u) * STRIP + bt;
                    if (u < nown && col < N) {
                        float yv = av[u];
                        if (p.rows) yv += box_take(cw[u], p.ws_cbox + col, err, boxtag, true);
                        p.ws_acc[col] = 0.f;
                        if (u == lidx) last_others = yv;
                        else store_y(col, yv);
                    }
                }
                for (int i = FAST; i < nown; ++i) {
                    const int col = (so0 + i) * STRIP + bt;
                    if (col < N) {
                        float yv = __ldcg(p.ws_acc + col);
                        if (p.rows) yv += box_take(0ull, p.ws_cbox + col, err, boxtag, true);
                        p.ws_acc[col] = 0.f;
                        if (i == lidx) last_others = yv;
                        else store_y(col, yv);
                    }
                }
#pragma unroll
                for (int h = 0; h < 3; ++h)
                    if (h < nh) last_others += box_take(hw[h], p.ws_hbox + ((size_t)blockIdx.x + 1 + h) * STRIP + bt, err, boxtag, true);
                for (int h = 3; h < nh; ++h) last_others += box_take(0ull, p.ws_hbox + ((size_t)blockIdx.x + 1 + h) * STRIP + bt, err, boxtag, true);
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
