// lutgemv_v2.cuh - round-2 kernel: ONE fat CTA per SM, TMA-fed weight ring, one strip table at a time, optional fp16 pair tables.
// Included by lutgemv_kernels.cu inside its anonymous namespace (shares the PTX helpers defined there).
//
// Same job as lutgemv_kernel (v1): y[c] (+)= sum_k LUT[c][idx(k,c)] * x[k]  (+ CSR outliers + topX dense rows), replacing
// squeezellm/quant_cuda_kernel.cu:741-880 (LUT GEMV), :1040-1059 (SPMV_ATOMIC), :1092-1123 (DenseMatVecKernel).
// What changed and why (measurements: profiles/r02_*; [*] = measured on an intermediate build in the first session of round 2, whose container and raw logs were lost):
//   * grid = #SMs, 23 warps per CTA: 16 consumer warps, 1 TMA producer warp, 2 table-builder warps, 4 sparse warps.  v1 ran 3 CTAs x 8 consumer warps per SM;
//     two thirds of its executed instructions on a 4096x4096 layer were per-CTA prologue/epilogue bookkeeping (ncu, r01).
//   * weights arrive as 2-D TMA boxes (cp.async.bulk.tensor.2d): stages are aligned to the CTA's strip segments, so a full stage is
//     ONE box of 64 columns x 32 units (8 KB / 24 KB) issued by one lane of the producer warp; the ragged last stage of a segment
//     goes out as 2-unit boxes, one per lane (second tensor map), so nothing is over-fetched.  (First attempt: one 256-byte
//     cp.async.bulk per row - 2.5x SLOWER than v1, the TMA unit retires a small copy every ~80 clocks: first session of round 2, log lost [*].)
//     Nothing of this touches the LSU pipe, which bounds the gather loop (32 gathers + 4 word reads + 2-4 x reads per 1024 weights).
//   * the CTA walks its strips one after the other with TWO table buffers in shared memory: the builder warps prepare the next strip's
//     table in the background and move each finished strip's sums (red.shared.add by the consumer warps) on to global memory; the
//     consumers never meet a CTA-wide barrier inside the loop.  Only two tables resident is what makes room for the
//   * fp16 PAIR table (MODE 1): entry [a | b<<BITS] = half2(LUT[a], LUT[b]) - one PRMT-built (4-bit: the byte of the packed word IS
//     the entry number) shared-memory lookup serves TWO weights, and the products go through fma.rn.f32.f16 (SASS FHFMA:
//     fp16 x fp16 -> fp32 accumulate, exact products).  Measured 48.9 (49.5 on another box) weights/clk/SM against 27.0 for the exact fp32 table
//     (tests/perf/microbench2.cu).  The only rounding is LUT -> fp16 (2^-11 relative per centroid); it needs fp16 x.
//   * x is staged whole (as fp16 when it is given as fp16), the sparse warps gather it from shared memory.
#pragma once

#ifndef SQLLM_CSR_LOCAL
#define SQLLM_CSR_LOCAL 0   // 1: a CTA sums the CSR rows of the strips it owns, in shared memory; 0: rows spread evenly over all CTAs (see sparse2)
#endif
#ifndef SQLLM_BOX
#define SQLLM_BOX 1         // 1: contributions to another CTA's strip travel as self-validating {value, valid} words (see "mailboxes" below); 0: red.add + flags
#endif
#ifndef SQLLM_V2_PIPE
#define SQLLM_V2_PIPE 0   // 1: two-stage ping-pong word fetch in the consumer loop (measured, not faster: see the loop)
#endif

namespace v2 {

constexpr int NWC = 16;                         // consumer warps
constexpr int WARP_PROD = NWC;                  // TMA producer
constexpr int WARP_BLD = NWC + 1;               // first of the 2 table-builder warps
constexpr int NBW = 2;
constexpr int WARP_SP = WARP_BLD + NBW;         // first of the sparse warps
constexpr int NSPW = 4;                         // sparse warps: their shared-memory accesses queue behind the consumers' (the LSU pipe is
                                                // saturated), so outlier work is latency-bound per warp - four warps, four times the progress
constexpr int THREADS2 = (NWC + 1 + NBW + NSPW) * 32;  // 736 (641..768 threads: 80 registers per thread)
constexpr int NCT = NWC * 32;                   // consumer threads
constexpr int NBT = NBW * 32;                   // builder threads: one per column slot of a strip
constexpr int SU2 = 2 * NWC;                    // units per stage (one pair per consumer warp)
constexpr int MAXD = 16;                        // ring depth limit (mbarrier slots)
constexpr int SP_CH = 256;                      // sparse warp: non-zeros per staged chunk (two chunk buffers per warp)
constexpr int SP_ROWS = 255;                    // owned rows whose pointers / sums are kept in shared memory (CTA-wide)
constexpr int SP_BYTES = 2 * SP_CH * 8;         // per sparse warp: 2 x (cols + vals) = 4 KB
constexpr int HYB_R2 = 11;
static_assert(NBT == STRIP, "one builder thread per column of a strip");
// shared-memory carve-up, offsets from a 128-byte aligned base
constexpr int OFF_BAR = 0;                      // full[s] at +8s, empty[s] at +128+8s, tfull[b] at +256+8b, tfree[b] at +272+8b
constexpr int OFF_SACC = 512;                   // float [2][64]: per-strip sums of the consumer warps (red.shared.add), ping-pong by strip parity
constexpr int OFF_CSR = OFF_SACC + 2 * STRIP * 4;            // per sparse warp: 2 x {int cols[SP_CH], float vals[SP_CH]}
constexpr int OFF_SROW = OFF_CSR + NSPW * SP_BYTES;          // CTA-wide: int srow[SP_ROWS + 1], float srowacc[SP_ROWS + 1]
constexpr int OFF_LUT = OFF_SROW + 2 * (SP_ROWS + 1) * 4;        // 2 x raw fp32 LUT rows of a strip (64 columns x 16 values), staged ahead by the builders
constexpr int LUTBUF = STRIP * (16 * 4 + 16);                // a strip's raw LUT rows, one per column SLOT, row stride L*4 + 16 bytes (see lut_prefetch)
constexpr int OFF_FIN = OFF_LUT + 2 * LUTBUF;                // finishers: float sown[8][64] (own strip sums), int sdj[128], float sden[128], int sdn (dense rows)
constexpr int OFF_X = OFF_FIN + 4096;                        // x (fp16 or fp32), then [ring stages][table 0][table 1][ring stages]
static_assert(OFF_X % 128 == 0, "x must stay 128-byte aligned");

struct P2 {
    const uint32_t *qw;
    const float *lut;
    const void *x;
    void *out;           // accumulate mode: float* mul ; fused: fp32/fp16 y
    const float *bias;
    const int *rows, *cols;
    const float *vals;
    const float *full_rows;
    const int *fri;
    int topX, K, N, R, T, chunk, nstage;
    unsigned smem_raw;   // shared-window address of the dynamic shared memory the host planned the carve-up for (checked by the kernel)
    int y_is_half;
    int hc, hrows, csr_rpc, csr_al16;
    float *ws_acc;       // fused: [N] fp32 accumulator, zero between launches
    unsigned long long *ws_hbox;  // fused, SQLLM_BOX: [grid][64] partial strip sums of a CTA's first segment when the strip starts in an earlier CTA
    unsigned long long *ws_cbox;  // fused, SQLLM_BOX: [N] outlier (CSR) row sums; both {float value, u32 valid}, all zero between launches
    unsigned long long *ws_dbox;  // sequence kernel: [topX][MAX_GRID_V2] dense-row partial sums, one tagged word per (dense row, contributing CTA)
    int *ws_cnt;         // fused: [16] error flag (a bounded wait gave up; sqllm_workspace_error), [64 + s] arrivals on strip s (zero between launches)
    int strips;          // output strips of 64 columns
    int nown_ctas;       // exchange: CTAs that own at least one strip (each announces itself once on every rank)
    int xw_world, xw_rank, xw_members, xw_nfull;
    const unsigned long long *xw_base;
    unsigned long long xw_out_off, xw_flag_off, xw_state_off, xw_err_off;
    int sp_w0; // 1: sparse warp 0 takes half a share of the CSR rows next to the dense rows (see sparse2)
    int l2pf;  // 1: prefetch the CTA's chunk of weights into L2 at entry
    unsigned long long *trace;
};

template <int BITS, int MODE>
struct C2 {
    static constexpr int L = 1 << BITS;
    static constexpr int ENT = MODE ? L * L : L;             // table rows
    static constexpr int TAB = ENT * STRIP * 4;              // 4 KB / 2 KB exact, 64 KB / 16 KB pair
    static constexpr int ROWS = BITS == 4 ? 1 : 3;           // packed rows per unit
    static constexpr int XU = BITS == 4 ? 8 : 32;            // inputs per unit
    static constexpr int UNIT = ROWS * STRIP * 4;            // bytes of one unit in a stage
    static constexpr int STAGE = SU2 * UNIT;                 // 8 KB / 24 KB
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// acc_lo += lo16(w) * lo16(x) ; acc_hi += hi16(w) * hi16(x)   (fp16 x fp16 products are exact in fp32; SASS: FHFMA with .H0/.H1 selectors)
__device__ __forceinline__ void fhfma2(float &acc_lo, float &acc_hi, uint32_t w, uint32_t x) {
    asm("{\n.reg .b16 wl, wh, xl, xh;\nmov.b32 {wl, wh}, %2;\nmov.b32 {xl, xh}, %3;\nfma.rn.f32.f16 %0, wl, xl, %0;\nfma.rn.f32.f16 %1, wh, xh, %1;\n}"
        : "+f"(acc_lo), "+f"(acc_hi) : "r"(w), "r"(x));
}
__device__ __forceinline__ uint64_t half2_to_f32x2(uint32_t h) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&h));
    return pack2(f.x, f.y);
}
__device__ __forceinline__ uint32_t f32x2_to_half2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Table build, by the NCT consumer threads.  Column c of the strip lives in slot ((c & 3) << 4) | (c >> 2): lane i16 of a half-warp
// owns columns 4*i16..4*i16+3, i.e. slots i16, 16+i16, 32+i16, 48+i16, so at every step the 32 lanes of a warp hit 32 different banks.
//   exact : row v, slot s  = LUT[c][v]                                  (fp32)
//   pair  : row a | b<<BITS = half2(lo = LUT[c][a], hi = LUT[c][b])      (a: the even input of the pair, b: the odd one)
// Columns past N read as zeros (ragged last strip: whatever index the stale stage bytes hold, the product is 0).
// ---------------------------------------------------------------------------------------------------------------------------
// Tables are built by the two builder warps (thread = slot = one column of the strip) from the strip's raw LUT rows, which they stage
// in shared memory first (lut_prefetch, zero-filled past N): a build never waits on global memory, and it runs in the background -
// the consumers only ever wait on an mbarrier that is normally long complete.  (First v2 build: the consumers rebuilt the table
// themselves at every strip boundary; ncu showed 24 % of all stall samples on the LUT loads and a 0.8 / 2.3 us bubble per switch.)
// ---------------------------------------------------------------------------------------------------------------------------
// Mailboxes (fused mode, SQLLM_BOX).  A sum that one CTA produces and another CTA needs - the partial strip sums of a CTA whose first
// segment continues a strip started earlier, and every outlier row sum (the CSR rows are spread evenly over the CTAs) - is written
// ONCE, as one 64-bit word {fp32 value, valid = 1}, and the strip's owner polls that word.  A 64-bit access is single-copy atomic, so
// the value needs no fence and no separate flag: the word validates itself.  Before: red.add into the accumulator, MEMBAR, flag
// increment on the producer; acquire-poll of the flag, then a second round trip for the accumulator on the owner - three L2 round
// trips (~2.5 us under load) between the last weight of a launch and its last y, now one.  The owner writes the word back to zero
// (next launch's producers only write after the grid dependency resolved).  The dense-row sums travel the same way since the second
// half of round 2: one word per (contributing CTA, dense row) - rows that feed the same channel are combined in the contributing warp -
// summed by the owner of the channel's strip; and the owner's OWN strip sums never leave the SM (sown, a shared-memory row per owned
// strip).  Before: red.add into the global accumulator + MEMBAR.GPU + one announcement per strip on the producers, an acquire-poll and a
// load of the accumulator behind it on the owner (7B w4-s45: 519.7 -> 547.0 tokens/s, profiles/r02_bench_llama7b_w4_s45.json).
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// `tag` is what makes a word valid: 1 for the one-GEMV kernel (the owner writes the word back to zero), a value unique to the (token, GEMV)
// for the sequence kernel (lutgemv_seq.cuh), where words are never reset (keep = true): a stale word simply carries an older tag.
__device__ __forceinline__ unsigned long long box_word(float v, uint32_t tag = 1u) { return ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v); }
// value of a mailbox word read speculatively as `w`; spins (bounded: 2 s, then the workspace error word) until it is valid; clears it
__device__ __forceinline__ float box_take(unsigned long long w, unsigned long long *ptr, int *err, const uint32_t tag = 1u, const bool keep = false) {
    if ((uint32_t)(w >> 32) != tag) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do {
            w = ld_relaxed_u64(ptr);
            if ((uint32_t)(w >> 32) == tag) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 2000000000ull) { *reinterpret_cast<volatile int *>(err) = 1; break; }
        } while (true);
    }
    if (!keep) st_relaxed_u64(ptr, 0ull);
    return __uint_as_float((uint32_t)w);
}

// Staging layout: the row of column c sits at slot(c) * (L*4 + 16) bytes.  The builders read it with 16-byte loads, lane = slot: with the
// rows packed by column (stride 4 columns x 64 B = 256 B between adjacent lanes) every LDS.128 was an 8-way bank conflict, and the
// pair-table build re-reads the row for each of its 16 outer steps - ncu: 1.18 M of the fp16 kernel's 3.69 M shared-memory wavefronts
// were these conflicts (first session of round 2, log lost [*]), as much LSU time as a third of the gathers.  With an 80-byte
// (4-bit) / 48-byte (3-bit) stride the 8 lanes of each 128-byte phase cover all 32 banks.
template <int BITS>
__device__ __forceinline__ void lut_prefetch(const P2 &p, const uint32_t lutbuf, const int strip, const int bt) {
    constexpr int L = 1 << BITS;
    constexpr int N16 = STRIP * L / 4;  // 16-byte pieces: 256 (w4) / 128 (w3)
#pragma unroll
    for (int e = bt; e < N16; e += NBT) {
        const int c = e / (L / 4), q = e % (L / 4), col = strip * STRIP + c;
        const int slot = ((c & 3) << 4) | (c >> 2);
        const bool ok = col < p.N;
        cp_async16_clip(lutbuf + slot * (L * 4 + 16) + 16 * q, p.lut + (ok ? (size_t)strip * STRIP * L + 4 * (size_t)e : 0), ok ? 16 : 0);
    }
    cp_async_commit();
}

template <int BITS, int MODE>
__device__ __forceinline__ void build_table(const uint32_t tab, const uint32_t lutbuf, const int slot) {
    using C = C2<BITS, MODE>;
    const uint32_t row = lutbuf + slot * (C::L * 4 + 16);    // the staged LUT row of the column this slot serves (slot s <-> column (s & 15) * 4 + (s >> 4))
    const uint32_t a0 = tab + (slot << 2);                   // a warp's 32 lanes = 32 consecutive slots: conflict-free stores
    float v[C::L];
#pragma unroll
    for (int q = 0; q < C::L / 4; ++q) {
        const float4 f = lds_v4(row + 16 * q);
        v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
    }
    if constexpr (MODE == 0) {
#pragma unroll
        for (int e = 0; e < C::L; ++e) sts_u32(a0 + e * (STRIP * 4), __float_as_uint(v[e]));
    } else {
        uint32_t H[C::L / 2];  // H[m] = half2(LUT[2m], LUT[2m+1])
#pragma unroll
        for (int m = 0; m < C::L / 2; ++m) H[m] = f32x2_to_half2(v[2 * m], v[2 * m + 1]);
#pragma unroll
        for (int b = 0; b < C::L; ++b) {
#pragma unroll
            for (int a = 0; a < C::L; ++a) {
                // row a | b << BITS : low half <- LUT[a] = half (a & 1) of H[a >> 1], high half <- LUT[b] = half (b & 1) of H[b >> 1]
                const uint32_t sel = (uint32_t)(((5 + 2 * (b & 1)) << 12) | ((4 + 2 * (b & 1)) << 8) | ((2 * (a & 1) + 1) << 4) | (2 * (a & 1)));
                sts_u32(a0 + ((a | (b << BITS)) << 8), __byte_perm(H[a >> 1], H[b >> 1], sel));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Per-position math.  A lane holds 4 adjacent columns x one unit; slot addresses l[t] (column t ^ jsel of the lane) are fixed
// for the whole kernel because there is one table.  Accumulators: exact -> 4 packed (even k, odd k) fp32 pairs; pair -> 8 fp32.
// ---------------------------------------------------------------------------------------------------------------------------
struct Acc {
    uint64_t a[4];
    float f[8];
};

// What a lane pulls out of a stage for one position: its 16 bytes (x 3 rows for 3-bit) and - 4-bit only, 8 inputs - its slice of x.
// Fetched one stage AHEAD of the math (software pipeline): the LDS.128 latency hides behind the previous position's gathers.
template <int BITS, bool XH> struct Fetch;
template <> struct Fetch<4, true> { uint4 q, xh; };
template <> struct Fetch<4, false> { uint4 q; float4 xa, xb; };
template <bool XH> struct Fetch<3, XH> { uint4 a, b, c; };

template <int BITS, bool XH>
__device__ __forceinline__ void fetch2(Fetch<BITS, XH> &F, const uint32_t unit_addr, const uint32_t xaddr) {
    if constexpr (BITS == 4) {
        F.q = lds_u4(unit_addr);
        if constexpr (XH) F.xh = lds_u4(xaddr);
        else { F.xa = lds_v4(xaddr); F.xb = lds_v4(xaddr + 16); }
    } else {
        F.a = lds_u4(unit_addr);
        F.b = lds_u4(unit_addr + STRIP * 4);
        F.c = lds_u4(unit_addr + 2 * STRIP * 4);
    }
}

// 4-bit, exact fp32 table (v1's loop): E/O = even/odd nibbles in separate bytes, OR-ed with bits 12..15 of the table address.
template <bool XH>
__device__ __forceinline__ void consume4_exact(const Fetch<4, XH> &F, const int jsel, const uint32_t (&l)[4], const uint32_t segc, Acc &A) {
    uint64_t x01, x23, x45, x67;
    if constexpr (XH) {
        x01 = half2_to_f32x2(F.xh.x); x23 = half2_to_f32x2(F.xh.y); x45 = half2_to_f32x2(F.xh.z); x67 = half2_to_f32x2(F.xh.w);
    } else {
        x01 = pack2(F.xa.x, F.xa.y); x23 = pack2(F.xa.z, F.xa.w); x45 = pack2(F.xb.x, F.xb.y); x67 = pack2(F.xb.z, F.xb.w);
    }
    const uint4 q = F.q;
    const uint32_t w[4] = {jsel ? q.y : q.x, jsel ? q.x : q.y, jsel ? q.w : q.z, jsel ? q.z : q.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const uint32_t E = (w[t] & 0x0F0F0F0Fu) | segc;
        const uint32_t O = ((w[t] >> 4) & 0x0F0F0F0Fu) | segc;
        const float e0 = lds_f32(__byte_perm(E, l[t], 0x7604)), o0 = lds_f32(__byte_perm(O, l[t], 0x7604));
        const float e1 = lds_f32(__byte_perm(E, l[t], 0x7614)), o1 = lds_f32(__byte_perm(O, l[t], 0x7614));
        const float e2 = lds_f32(__byte_perm(E, l[t], 0x7624)), o2 = lds_f32(__byte_perm(O, l[t], 0x7624));
        const float e3 = lds_f32(__byte_perm(E, l[t], 0x7634)), o3 = lds_f32(__byte_perm(O, l[t], 0x7634));
        ffma2(A.a[t], pack2(e0, o0), x01);
        ffma2(A.a[t], pack2(e1, o1), x23);
        ffma2(A.a[t], pack2(e2, o2), x45);
        ffma2(A.a[t], pack2(e3, o3), x67);
    }
}
// 4-bit, fp16 pair table (64 KB aligned): byte j of a packed word = (index of input 2j) | (index of input 2j+1) << 4 = the entry number.
__device__ __forceinline__ void consume4_pair(const Fetch<4, true> &F, const int jsel, const uint32_t (&l)[4], Acc &A) {
    const uint4 q = F.q, xh = F.xh;  // halves (x0,x1) (x2,x3) (x4,x5) (x6,x7)
    const uint32_t w[4] = {jsel ? q.y : q.x, jsel ? q.x : q.y, jsel ? q.w : q.z, jsel ? q.z : q.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const uint32_t p0 = lds_u32(__byte_perm(w[t], l[t], 0x7604)), p1 = lds_u32(__byte_perm(w[t], l[t], 0x7614));
        const uint32_t p2 = lds_u32(__byte_perm(w[t], l[t], 0x7624)), p3 = lds_u32(__byte_perm(w[t], l[t], 0x7634));
        fhfma2(A.f[t], A.f[t + 4], p0, xh.x);
        fhfma2(A.f[t], A.f[t + 4], p1, xh.y);
        fhfma2(A.f[t], A.f[t + 4], p2, xh.z);
        fhfma2(A.f[t], A.f[t + 4], p3, xh.w);
    }
}
// 3-bit pair table (16 KB aligned): the three words of a group are a 96-bit stream of 32 3-bit fields (quant.py:185-203), so pair
// p = (input 2p, input 2p+1) is the 6-bit field at stream bit 6p.  Moved to bits 8..13 by one shift (a funnel shift for pairs 5, 10).
template <int SH>
__device__ __forceinline__ uint32_t fld8(uint32_t w) {
    if constexpr (SH < 8) return w << (8 - SH);
    else if constexpr (SH == 8) return w;
    else return w >> (SH - 8);
}
#define LKP(e, word_expr) e = lds_u32((((word_expr)) & 0x3F00u) | lsv)
__device__ __forceinline__ void consume3_pair_col(const uint32_t w0, const uint32_t w1, const uint32_t w2, const uint32_t lsv,
                                                  const uint32_t (&xh)[16], float &fa, float &fb) {
    uint32_t e;
    LKP(e, fld8<0>(w0));  fhfma2(fa, fb, e, xh[0]);
    LKP(e, fld8<6>(w0));  fhfma2(fa, fb, e, xh[1]);
    LKP(e, fld8<12>(w0)); fhfma2(fa, fb, e, xh[2]);
    LKP(e, fld8<18>(w0)); fhfma2(fa, fb, e, xh[3]);
    LKP(e, fld8<24>(w0)); fhfma2(fa, fb, e, xh[4]);
    LKP(e, __funnelshift_r(w0, w1, 22)); fhfma2(fa, fb, e, xh[5]);   // stream bits 30..35
    LKP(e, fld8<4>(w1));  fhfma2(fa, fb, e, xh[6]);
    LKP(e, fld8<10>(w1)); fhfma2(fa, fb, e, xh[7]);
    LKP(e, fld8<16>(w1)); fhfma2(fa, fb, e, xh[8]);
    LKP(e, fld8<22>(w1)); fhfma2(fa, fb, e, xh[9]);
    LKP(e, __funnelshift_r(w1, w2, 20)); fhfma2(fa, fb, e, xh[10]);  // stream bits 60..65
    LKP(e, fld8<2>(w2));  fhfma2(fa, fb, e, xh[11]);
    LKP(e, fld8<8>(w2));  fhfma2(fa, fb, e, xh[12]);
    LKP(e, fld8<14>(w2)); fhfma2(fa, fb, e, xh[13]);
    LKP(e, fld8<20>(w2)); fhfma2(fa, fb, e, xh[14]);
    LKP(e, fld8<26>(w2)); fhfma2(fa, fb, e, xh[15]);
}
#undef LKP

template <int BITS, int MODE, bool XH>
__device__ __forceinline__ void math2(const Fetch<BITS, XH> &F, const int jsel, const uint32_t (&l)[4], const uint32_t segc, const uint32_t xaddr, Acc &A) {
    if constexpr (BITS == 4) {
        if constexpr (MODE == 0) consume4_exact<XH>(F, jsel, l, segc, A);
        else consume4_pair(F, jsel, l, A);
    } else {
        const uint4 ga = F.a, gb = F.b, gc = F.c;
        const uint32_t a[4] = {jsel ? ga.y : ga.x, jsel ? ga.x : ga.y, jsel ? ga.w : ga.z, jsel ? ga.z : ga.w};
        const uint32_t b[4] = {jsel ? gb.y : gb.x, jsel ? gb.x : gb.y, jsel ? gb.w : gb.z, jsel ? gb.z : gb.w};
        const uint32_t c[4] = {jsel ? gc.y : gc.x, jsel ? gc.x : gc.y, jsel ? gc.w : gc.z, jsel ? gc.z : gc.w};
        if constexpr (MODE == 0) {
            uint64_t xp[16];
            if constexpr (XH) {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const uint4 h = lds_u4(xaddr + 16 * v);
                    xp[4 * v] = half2_to_f32x2(h.x); xp[4 * v + 1] = half2_to_f32x2(h.y);
                    xp[4 * v + 2] = half2_to_f32x2(h.z); xp[4 * v + 3] = half2_to_f32x2(h.w);
                }
            } else {
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    const float4 f = lds_v4(xaddr + 16 * v);
                    xp[2 * v] = pack2(f.x, f.y);
                    xp[2 * v + 1] = pack2(f.z, f.w);
                }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) consume3_col(a[t], b[t], c[t], l[t], xp, A.a[t]);
        } else {
            uint32_t xh[16];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const uint4 h = lds_u4(xaddr + 16 * v);
                xh[4 * v] = h.x; xh[4 * v + 1] = h.y; xh[4 * v + 2] = h.z; xh[4 * v + 3] = h.w;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) consume3_pair_col(a[t], b[t], c[t], l[t], xh, A.f[t], A.f[t + 4]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Sparse warp(s) (NSPW per CTA).  CSR rows (= output channels) are spread evenly over the CTAs (csr_rpc consecutive rows each; split
// between the CTA's sparse warps if there are several); warp 0 also takes the CTA's k-slice of the topX dense rows.  Everything static (row
// pointers, the first cols/vals group, dense-row values) is requested before the x barrier; x comes from shared memory.
// Per-row sums are taken in storage order; results go out as one red.add per row.
// ---------------------------------------------------------------------------------------------------------------------------
template <bool XH>
__device__ __forceinline__ float xs_load(const uint32_t xs_u32, const int k) {
    if constexpr (XH) {
        unsigned short h;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(xs_u32 + 2 * k));
        return __half2float(__ushort_as_half(h));
    } else {
        return lds_f32(xs_u32 + 4 * k);
    }
}

template <bool XH, bool FUSED>
__device__ __forceinline__ void sparse2(const P2 &p, unsigned char *sm, const uint32_t base, const int spw, const int lane, float *acc_out, const uint32_t boxtag = 1u) {
    const int N = p.N;
    constexpr bool BOX = FUSED && SQLLM_BOX && !SQLLM_CSR_LOCAL;  // cross-CTA sums travel as mailbox words
    const uint32_t xs_u32 = base + OFF_X;
    const uint32_t stage_u32 = base + OFF_CSR + spw * SP_BYTES;       // this warp's chunk buffer b: cols at +b*SP_CH*8, vals SP_CH*4 further
    int *srow = reinterpret_cast<int *>(sm + OFF_SROW);               // [SP_ROWS + 1] row pointers of the CTA's rows (FUSED, when they fit)
    float *srowacc = reinterpret_cast<float *>(srow + SP_ROWS + 1);   // [SP_ROWS + 1] their outlier sums, picked up by the builders

    // ---------------- phase A: static data ----------------
    float hfr[HYB_R2];
    const bool hyb_on = p.full_rows && spw == 0 && (int)blockIdx.x < p.hc;
    const bool hyb_multi = hyb_on && p.topX <= 32;  // 32/topX k-rows per warp step: lane = (row slot, column j)
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
    // FUSED: the CSR rows of the strips this CTA OWNS (those that start in its range - it also finishes them), so that outlier sums
    // never leave the CTA: they go into a shared-memory row accumulator that the builders add when they write y.  (Earlier v2 builds
    // spread the rows evenly over all CTAs and announced them on the strips' flags: a MEMBAR.GPU per announcement, and every strip owner
    // waited for some other CTA's sparse warp - the critical path of the kernel on layers with outliers.)
    // Accumulate mode (the 12 symbols, no owner): rows spread evenly over the CTAs, one red.add per row into `mul`.
    // The CTA's rows [ca, cb) are split over its NSPW sparse warps (warp 0, which also has the dense rows, takes half a share).
    int ca = 0, cb = 0;
    if (p.rows) {
        if constexpr (FUSED && SQLLM_CSR_LOCAL) {
            const long long cb0 = (long long)blockIdx.x * p.chunk, cb1 = min((long long)p.T, cb0 + p.chunk);
            ca = min(N, (int)((cb0 + p.R - 1) / p.R) * STRIP);
            cb = min(N, (int)((cb1 + p.R - 1) / p.R) * STRIP);
        } else {
            ca = min(N, (int)blockIdx.x * p.csr_rpc);
            cb = min(N, ca + p.csr_rpc);
        }
    }
    int r, rb;
    {
        // warp 0 also has the dense rows (atomics, a fence, the announcement: 2-3 us after x arrives).  sp_w0 = 0 (default): it then takes no CSR
        // rows at all - the outlier sums are what the owners of ALL strips wait for, they must not queue behind that fence; 1: half a share (round-2
        // first build; A/B with SQLLM_SP_W0=1)
        const int tot = cb - ca, w0 = p.full_rows ? (p.sp_w0 ? tot / (2 * NSPW) : 0) : tot / NSPW, rest = tot - w0;
        r = spw == 0 ? ca : ca + w0 + (int)((long long)rest * (spw - 1) / (NSPW - 1));
        rb = spw == 0 ? ca + w0 : ca + w0 + (int)((long long)rest * spw / (NSPW - 1));
    }
    const int nr = rb - r;
    const bool rows_in_smem = cb - ca <= SP_ROWS;                       // the CTA's row pointers live in shared memory
    const bool local_sums = FUSED && SQLLM_CSR_LOCAL && rows_in_smem;  // ... and so do the row sums (same rule in the builders)
    // (BOX: row sums are published as mailbox words when the warp is through)
    const bool smem_sums = local_sums || (BOX && rows_in_smem);
    auto emit = [&](int row, float v) {
        if (smem_sums) srowacc[row - ca] += v;  // plain read-modify-write: a row belongs to this warp alone and its pieces arrive one after the other
        else if (BOX) atomicAdd(reinterpret_cast<float *>(p.ws_cbox + row), v);  // (more rows per CTA than the shared accumulator holds: value half of the word)
        else atomicAdd(acc_out + row, v);
    };
    int e_lo = 0, e_hi = 0;
    if (nr > 0) {
        if (rows_in_smem)
            for (int i = lane; i <= nr; i += 32) {
                srow[r - ca + i] = __ldg(p.rows + r + i);   // (the boundary entry is written by both neighbours, with the same value)
                if (i < nr) srowacc[r - ca + i] = 0.f;
            }
        e_lo = __ldg(p.rows + r);
        e_hi = __ldg(p.rows + rb);
    }
    // This warp's non-zeros [e_lo, e_hi) go to L2 right away (a CTA has ~5-25 KB of outliers and only two small chunks per warp fit the
    // staging buffers: under a saturated HBM every further chunk would otherwise cost a DRAM round trip of 2-3 us), and are then walked
    // in chunks of SP_CH staged with 16-byte cp.async into two buffers.  Shared-memory accesses of these warps queue behind the
    // consumers' (the LSU pipe is what bounds the kernel), i.e. each dependent access costs hundreds of cycles: that is why there are
    // NSPW warps with independent chains (one warp: 14 us for 3200 non-zeros) and why every step below batches independent loads.
    for (int e = (e_lo & ~31) + 32 * lane; e < e_hi; e += 32 * 32) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.cols + e));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.vals + e));
    }
    const bool al16 = p.csr_al16 != 0;
    const int c_lo = al16 ? (e_lo & ~3) : e_lo;  // chunk grid: 16-byte aligned element index when the arrays are
    auto stage_chunk = [&](int ci) {              // chunk ci covers elements [c_lo + ci*SP_CH, +SP_CH) -> buffer ci & 1
        const int t0 = c_lo + ci * SP_CH;
        const uint32_t dc = stage_u32 + (ci & 1) * (SP_CH * 8), dv = dc + SP_CH * 4;
        if (t0 < e_hi) {
            if (al16) {
                for (int e = 4 * lane; e < SP_CH; e += 128) {
                    const int nb = min(16, 4 * (e_hi - (t0 + e)));  // the last quad is clipped with the src-size operand: nothing past e_hi is read
                    if (nb > 0) {
                        cp_async16_clip(dc + 4 * e, p.cols + t0 + e, nb);
                        cp_async16_clip(dv + 4 * e, p.vals + t0 + e, nb);
                    }
                }
            } else {
                for (int e = lane; e < SP_CH && t0 + e < e_hi; e += 32) {
                    cp_async4(dc + 4 * e, p.cols + t0 + e);
                    cp_async4(dv + 4 * e, p.vals + t0 + e);
                }
            }
        }
        cp_async_commit();  // exactly one group per chunk, empty or not
    };
    stage_chunk(0);
    stage_chunk(1);

    named_bar_sync(2, NCT + NSPW * 32);  // x is in shared memory (and the previous kernel has completed: the consumers waited on it)

    // ---------------- phase B ----------------
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
                for (int sl = 1; sl < nsl; ++sl) {  // fold the row slots onto slot 0 in fixed order
                    const float v = __shfl_sync(0xffffffffu, a, (hj + sl * p.topX) & 31);
                    if (rs == 0) a += v;
                }
                if (rs != 0) j = p.topX;  // only slot 0 publishes
                if constexpr (BOX) {
                    // dense rows that feed the SAME channel travel as one word (the first of them carries the sum): the owner of that channel's
                    // strip then reads hc words, not hc per row - a checkpoint without dense rows loads as topX rows on channel 0 (llama.py:182)
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
            if constexpr (BOX) {
                // this CTA's part of dense row j: one self-validating word, no RED + fence + flag (the owner of the channel's strip sums the hc words of
                // row j, lutgemv2_kernel's finishers); before: atomics into the accumulator, a MEMBAR.GPU and an announcement per strip - the fence
                // alone held this warp for 1-2 us, and with it every owner that waited for the announcement
                if (j < p.topX) st_relaxed_u64(p.ws_dbox + (size_t)j * MAX_GRID_V2 + blockIdx.x, box_word(a));
            } else if (j < p.topX) {
                const int c = __ldg(p.fri + j);
                if (c >= 0 && c < N) atomicAdd(acc_out + c, a);
            }
        }
        if constexpr (FUSED && SQLLM_CSR_LOCAL) {
            // announce this CTA's dense-row contributions now, early in the kernel: one fence, then a relaxed increment per distinct
            // strip that holds a dense-row channel (every one of the hc contributing CTAs does this; the owners expect hc arrivals)
            __threadfence();
            __syncwarp();
            int *flags = p.ws_cnt + 64;
            for (int j = lane; j < p.topX; j += 32) {
                const int c = __ldg(p.fri + j);
                if (c < 0 || c >= N) continue;
                bool seen = false;
                for (int j2 = 0; j2 < j; ++j2) {
                    const int c2 = __ldg(p.fri + j2);
                    seen |= (c2 >= 0 && c2 < N && c2 / STRIP == c / STRIP);
                }
                if (!seen) asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(flags + c / STRIP) : "memory");
            }
        }
    }
    int rcur = r;  // first row that still has non-zeros at or after the current chunk
    const int avg = nr > 0 ? (e_hi - e_lo) / nr : 0;
    const int LPR = avg > 32 ? 8 : avg > 12 ? 4 : avg > 5 ? 2 : 1;  // lanes per row (power of two)
    for (int ci = 0; c_lo + ci * SP_CH < e_hi; ++ci) {
        const int t0 = c_lo + ci * SP_CH, t1 = min(t0 + SP_CH, e_hi), f0 = max(t0, e_lo);
        const int *scols = reinterpret_cast<const int *>(sm + OFF_CSR + spw * SP_BYTES + (ci & 1) * (SP_CH * 8));
        float *svals = reinterpret_cast<float *>(sm + OFF_CSR + spw * SP_BYTES + (ci & 1) * (SP_CH * 8) + SP_CH * 4);
        cp_async_wait_pending<1>();  // all but the newest group have landed: this chunk is there
        __syncwarp();
        // products in place: all (col, val) loads of a trip first, then the x gathers, then the stores
        for (int e0 = f0 - t0 + lane; e0 < t1 - t0; e0 += 256) {
            int cc[8];
            float xv[8], vv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool ok = e0 + 32 * j < t1 - t0;
                cc[j] = ok ? scols[e0 + 32 * j] : 0;
                vv[j] = ok ? svals[e0 + 32 * j] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) xv[j] = xs_load<XH>(xs_u32, cc[j]);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (e0 + 32 * j < t1 - t0) svals[e0 + 32 * j] = vv[j] * xv[j];
        }
        __syncwarp();
        // rows that intersect [f0, t1), 32 / LPR at a time from rcur: LPR lanes share a row (lane part p takes elements p, p + LPR, ...
        // of the row's piece, four independent loads per trip), their partial sums are folded with shuffles and the row's piece is
        // added to its sum (a row that straddles chunks arrives in pieces).  LPR grows with the average row length: a dependent
        // shared-memory trip costs this warp hundreds of cycles, a 50-element row summed by one lane would be 13 of them.
        while (true) {
            const int row = rcur + lane / LPR, part = lane & (LPR - 1);
            int a0 = e_hi, a1 = e_hi;
            if (row < rb) {
                a0 = rows_in_smem ? srow[row - ca] : __ldg(p.rows + row);
                a1 = rows_in_smem ? srow[row - ca + 1] : __ldg(p.rows + row + 1);
            }
            const bool inter = row < rb && a0 < t1;
            const int s0 = max(a0, f0) - t0, s1 = min(a1, t1) - t0;
            const int n = inter ? s1 - s0 : 0;
            float ea = 0.f, eb = 0.f, ec = 0.f, ed = 0.f;
            if (n > 0 && n <= 64 * LPR) {
                for (int e = s0 + part; e < s1; e += 4 * LPR) {  // four independent (predicated) loads per trip
                    const float q0 = svals[e];
                    const float q1 = e + LPR < s1 ? svals[e + LPR] : 0.f;
                    const float q2 = e + 2 * LPR < s1 ? svals[e + 2 * LPR] : 0.f;
                    const float q3 = e + 3 * LPR < s1 ? svals[e + 3 * LPR] : 0.f;
                    ea += q0; eb += q1; ec += q2; ed += q3;
                }
            }
            float tot = (ea + eb) + (ec + ed);
            for (int d = 1; d < LPR; d <<= 1) tot += __shfl_xor_sync(0xffffffffu, tot, d);
            if (part == 0 && n > 0 && n <= 64 * LPR) emit(row, tot);
            unsigned longm = __ballot_sync(0xffffffffu, part == 0 && n > 64 * LPR);  // very long pieces: the whole warp on each
            while (longm) {
                const int i = __ffs(longm) - 1;
                longm &= longm - 1;
                const int b0 = __shfl_sync(0xffffffffu, s0, i), b1 = __shfl_sync(0xffffffffu, s1, i);
                float a = 0.f;
                for (int e = b0 + lane; e < b1; e += 32) a += svals[e];
                a = warp_sum(a);
                if (lane == 0) emit(rcur + i / LPR, a);
            }
            const int ndone = __popc(__ballot_sync(0xffffffffu, part == 0 && row < rb && a1 <= t1));  // rows are sorted: a prefix
            rcur += ndone;
            if (ndone < 32 / LPR || rcur >= rb) break;  // the last row of the pass still has non-zeros beyond this chunk (or no rows are left)
        }
        __syncwarp();
        stage_chunk(ci + 2);  // this buffer is free again
    }
    if constexpr (FUSED && SQLLM_CSR_LOCAL) {
        if (!local_sums && cb > ca) __threadfence();  // (more owned rows than the row accumulator holds: sums went to global memory)
        named_bar_sync(4, (NSPW + NBW) * 32);         // hand-over to the builders: row sums complete
    } else if constexpr (BOX) {
        // every row of this warp's share gets its word, outlier-free rows included (the owner of the strip waits for all 64 of them)
        __syncwarp();
        if (smem_sums) {
            for (int i = lane; i < nr; i += 32) st_relaxed_u64(p.ws_cbox + r + i, box_word(srowacc[r - ca + i], boxtag));
        } else if (nr > 0) {
            __threadfence();  // the red.adds on the value halves come first
            for (int i = lane; i < nr; i += 32) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(reinterpret_cast<uint32_t *>(p.ws_cbox + r + i) + 1), "r"(1u) : "memory");
        }
    } else if constexpr (FUSED) {
        // Balanced mode: this CTA's outlier sums went to the global accumulator with red.add.  Announce them: all sparse warps meet,
        // then ONE fence (by warp 0) and a relaxed increment per strip the CTA's rows touch (whatever their nnz) and per distinct
        // strip that holds a dense-row channel - the owners of those strips expect exactly these arrivals.
        named_bar_sync(5, NSPW * 32);
        if (spw == 0) {
            __threadfence();
            __syncwarp();
            int *flags = p.ws_cnt + 64;
            if (cb > ca)
                for (int s = ca / STRIP + lane; s <= (cb - 1) / STRIP; s += 32)
                    asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(flags + s) : "memory");
            if (p.full_rows && (int)blockIdx.x < p.hc)
                for (int j = lane; j < p.topX; j += 32) {
                    const int c = __ldg(p.fri + j);
                    if (c < 0 || c >= N) continue;
                    bool seen = false;
                    for (int j2 = 0; j2 < j; ++j2) {
                        const int c2 = __ldg(p.fri + j2);
                        seen |= (c2 >= 0 && c2 < N && c2 / STRIP == c / STRIP);
                    }
                    if (!seen) asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(flags + c / STRIP) : "memory");
                }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// The kernel.  MODE 0: exact fp32 table ; MODE 1: fp16 pair table (needs XH).  XH: x is given (and staged) as fp16.
// ---------------------------------------------------------------------------------------------------------------------------
template <int BITS, int MODE, bool XH, bool FUSED>
__global__ void __launch_bounds__(THREADS2, 1) lutgemv2_kernel(const P2 p, const __grid_constant__ CUtensorMap tm_big,
                                                               const __grid_constant__ CUtensorMap tm_small) {
    using C = C2<BITS, MODE>;
    static_assert(MODE == 0 || XH, "the pair table multiplies fp16 by fp16: it needs fp16 x");
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    if (raw != p.smem_raw) __trap();  // the host planned the carve-up (table alignment) for another window address
    uint32_t base = (raw + 127u) & ~127u;
    asm volatile("mov.u32 %0, %0;" : "+r"(base));  // pinned: keeps the compiler from rematerialising it through S2UR
    unsigned char *sm = smem_raw + (base - raw);
    const uint32_t bar_u32 = base + OFF_BAR;
    const uint32_t xs_u32 = base + OFF_X;
    const uint32_t lo_base = xs_u32 + (uint32_t)((p.K * (XH ? 2 : 4) + 127) & ~127);
    const uint32_t tab0 = (lo_base + (uint32_t)C::TAB - 1u) & ~((uint32_t)C::TAB - 1u);  // two tables, each aligned to its own size
    const int nstage = p.nstage;
    const int n_lo = min(nstage, (int)((tab0 - lo_base) / C::STAGE));       // ring stages that fit below the aligned tables
    const uint32_t hi_base = tab0 + 2 * C::TAB - (uint32_t)n_lo * C::STAGE; // stage s >= n_lo lives at hi_base + s * STAGE

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, R = p.R;
    const int g0 = min((int)blockIdx.x * p.chunk, p.T), g1 = min(g0 + p.chunk, p.T);
    const int len = g1 - g0;
    const int s0 = g0 / R, r0 = g0 - s0 * R;
    const int nseg = len > 0 ? (g1 - 1) / R - s0 + 1 : 0;

    TRACE(0, tid == 0);
    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) {
            mbar_init(bar_u32 + 8 * s, 1);          // full: the producer's arrive.expect_tx
            mbar_init(bar_u32 + 128 + 8 * s, NWC);  // empty: one arrive per consumer warp
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_u32 + 256 + 8 * b, NBW);  // tfull: table b is built (one arrive per builder warp)
            mbar_init(bar_u32 + 272 + 8 * b, NWC);  // tfree: every consumer warp is done with table b and has deposited its strip sums
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_launch_dependents();
    if (warp == WARP_PROD) {
        // Cold-start latencies (a decode step finds none of a layer's arrays in L2): ask for the two tensor-map descriptors now, and let
        // every CTA pull a 1/grid slice of the layer's look-up table and row pointers into L2 - the CTAs of a launch start up to ~4 us
        // apart (they inherit their SM from the previous kernel's CTAs), so the late ones find what the early ones asked for.
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_big)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_small)) : "memory");
        }
        if (p.l2pf) {
            // the CTA's whole chunk of weights -> L2, one box per lane and instruction, before the dependency wait: the ring then only has
            // to cover L2 latency, and under PDL these requests overlap the previous kernel's tail
            for (int seg = 0; seg < nseg; ++seg) {
                const int sstart = seg == 0 ? 0 : seg * R - r0, send = min(len, (seg + 1) * R - r0);
                const int u0 = r0 + sstart - seg * R, col0 = (s0 + seg) * STRIP;
                for (int u = lane * SU2; u < send - sstart; u += 32 * SU2)
                    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(&tm_big)),
                                 "r"(col0), "r"((u0 + u) * C::ROWS) : "memory");
            }
        }
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
    __syncthreads();
    float *const acc_out = FUSED ? p.ws_acc : reinterpret_cast<float *>(p.out);

    if (warp == WARP_PROD) {
        // =========================== TMA producer: weights never depend on the previous kernel ===========================
        // Stages never straddle a strip: segment `seg` (the CTA's part of strip s0+seg) is cut into stages of SU2 units from its own start.
        // (Also tried: cp.async.bulk.prefetch.tensor into L2 16-24 stages ahead of the ring - 5-10 % slower on every shape, like the
        // L2 prefetch experiment of round 1: first session of round 2, log lost [*].)
        const uint64_t pol = l2_evict_first_policy();
        int slot = 0;
        uint32_t ph = 0;
        bool refill = false;
        for (int seg = 0; seg < nseg; ++seg) {
            const int sstart = seg == 0 ? 0 : seg * R - r0, send = min(len, (seg + 1) * R - r0);
            const int seglen = send - sstart;
            const int u0 = r0 + sstart - seg * R;  // unit-in-strip of the segment's first unit
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
                    if (lane == 0) tma_tile2d_g2s(dst, &tm_big, col0, row0, full, pol);
                } else if (2 * lane < nu) {
                    tma_tile2d_g2s(dst + lane * 2 * C::UNIT, &tm_small, col0, row0 + lane * 2 * C::ROWS, full, pol);
                }
                if (++slot == nstage) {
                    slot = 0;
                    if (refill) ph ^= 1u;
                    refill = true;
                }
            }
        }
    } else if (warp >= WARP_SP) {
        sparse2<XH, FUSED>(p, sm, base, warp - WARP_SP, lane, acc_out);
        TRACE(10, lane == 0);
    } else if (warp >= WARP_BLD) {
        // =========================== table builders (2 warps, thread = column slot) ===========================
        // Segment s uses table buffer s & 1 and strip accumulator s & 1.  For every segment: wait until the consumers are done with the
        // buffer's previous user (segment s-2), flush that segment's sums to global memory, build the new table, signal tfull.
        const int bt = tid - WARP_BLD * 32;
        const uint32_t lutbuf = base + OFF_LUT, sacc = base + OFF_SACC;
        sts_u32(sacc + 4 * bt, 0u);
        sts_u32(sacc + 4 * (STRIP + bt), 0u);
        int *const flags = p.ws_cnt + 64;
        bool dep_ok = false;
        auto dep_wait = [&]() { if (!dep_ok) { pdl_wait(); dep_ok = true; } };  // builders never read x: they only need this before writing
        const uint32_t sown = base + OFF_FIN;   // float [SOWN][64]
        constexpr int SOWN = 8;
        auto flush = [&](int s) {  // strip sums of segment s -> one red.add per column, accumulator back to zero
            const uint32_t a = sacc + 4 * ((s & 1) * STRIP + bt);
            const float v = lds_f32(a);
            sts_u32(a, 0u);
            const int strip = s0 + s, col = strip * STRIP + bt;
            if constexpr (FUSED && SQLLM_BOX && !SQLLM_CSR_LOCAL) {
                // A strip is finished (converted to y) by the CTA in whose range it STARTS; a CTA that only holds a later part of it
                // (that can only be its first segment) hands its sums to that owner through its mailbox row.
                // ... or to ourselves: then they never leave the SM (sown[u][column] for the u-th owned strip; a strip inside one CTA is exactly
                // one segment).  Through the global accumulator (first round-2 build) that was a RED and, in the finishing pass, a load that had to
                // wait for it - two L2 round trips of 1-2.5 us each between the last weight and y when the last segment is short.
                if (s == 0 && r0 != 0) st_relaxed_u64(p.ws_hbox + (size_t)blockIdx.x * STRIP + bt, box_word(v));
                else if (s - (r0 != 0 ? 1 : 0) < SOWN) sts_u32(sown + 4u * (uint32_t)((s - (r0 != 0 ? 1 : 0)) * STRIP + bt), __float_as_uint(v));
                else if (col < N) atomicAdd(acc_out + col, v);
            } else {
                if (col < N) atomicAdd(acc_out + col, v);
                if constexpr (FUSED) {
                    // (flag protocol) ... announces its contribution on the strip's flag
                    if (s == 0 && r0 != 0) {
                        named_bar_sync(3, NBT);
                        if (bt == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(flags + strip) : "memory");
                    }
                }
            }
        };
        // The raw LUT rows of the first two strips are requested at once (one cold round trip, not two); later strips two ahead.
        if (nseg > 0) lut_prefetch<BITS>(p, lutbuf, s0, bt);
        if (nseg > 1) lut_prefetch<BITS>(p, lutbuf + LUTBUF, s0 + 1, bt);
        for (int s = 0; s < nseg; ++s) {
            const int b = s & 1;
            if (s >= 2) {
                dep_wait();  // from here on we write global memory: the previous kernel must be done
                mbar_wait(bar_u32 + 272 + 8 * b, (uint32_t)(((s >> 1) - 1) & 1));
                flush(s - 2);
            }
            if (s + 1 < nseg) cp_async_wait_pending<1>();  // all but the newest group (the next strip's rows) have landed
            else cp_async_wait_all();
            named_bar_sync(3, NBT);  // the strip's LUT rows are in shared memory, all of them
            build_table<BITS, MODE>(tab0 + b * C::TAB, lutbuf + b * LUTBUF, bt);
            named_bar_sync(3, NBT);  // both warps are done reading the rows (and writing the table)
            if (lane == 0) mbar_arrive(bar_u32 + 256 + 8 * b);
            if (s + 2 < nseg) lut_prefetch<BITS>(p, lutbuf + b * LUTBUF, s0 + s + 2, bt);  // this buffer's next strip: long there when it is needed
        }
        dep_wait();
        if constexpr (!FUSED) {
            for (int s = max(0, nseg - 2); s < nseg; ++s) {  // the last two segments
                mbar_wait(bar_u32 + 272 + 8 * (s & 1), (uint32_t)((s >> 1) & 1));
                flush(s);
            }
        } else {
            // ---- finish the strips this CTA owns (those that start in its range).  Every table is built, so the builders have time: while
            //      the consumers are still in the last segment they wait for the other contributors' announcements - dense CTAs holding a
            //      later part, the hc dense-row CTAs if a dense-row channel lies in it (the strips' CSR outliers are summed by this CTA's
            //      own sparse warp, in shared memory) - and fetch the accumulator.  What is left for the very end is y = accumulator + own last strip (+ bias): no round trip to
            //      L2 after the last weight, no grid-wide step; a CTA leaves as soon as its own strips are complete.
            //      Waits are bounded (2 s, then the workspace error word is set).
#if SQLLM_BOX && !SQLLM_CSR_LOCAL
            // ---- mailbox variant (shared with the sequence kernel, lutgemv_seq.cuh): own strip sums from shared memory, outlier sums and the later
            //      CTAs' first-segment sums from their mailbox words, dense rows from one tagged word per (contributing CTA, row).
            TRACE(12, bt == 0);
            const long long cb0 = (long long)blockIdx.x * p.chunk, cb1 = min((long long)p.T, cb0 + p.chunk);
            const int so0 = (int)((cb0 + R - 1) / R), so1 = (int)((cb1 + R - 1) / R), nown = so1 - so0;
            const bool last_owned = nseg >= 2 || r0 == 0;
            constexpr int FAST = 4;
            const int lidx = last_owned && nseg > 0 ? nown - 1 : -1;
            const int nh = lidx >= 0 ? (int)((((long long)(so1 - 1) + 1) * R - 1) / p.chunk) - (int)blockIdx.x : 0;
            // Dense rows whose channel lies in a strip we own: listed now (static), summed below.  sdj[i] = dense row, sden[i] = its sum.
            int *const err = p.ws_cnt + 16;
            int *const sdj = reinterpret_cast<int *>(sm + OFF_FIN + 2048);
            const uint32_t sden = base + OFF_FIN + 2048 + 4 * 128, sdn_a = base + OFF_FIN + 2048 + 8 * 128;
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
            for (int s = max(0, nseg - 2); s < nseg - 1; ++s) {
                mbar_wait(bar_u32 + 272 + 8 * (s & 1), (uint32_t)((s >> 1) & 1));
                flush(s);
            }
            TRACE(13, bt == 0);
            {
                unsigned long long cw[FAST], hw[3];
#pragma unroll
                for (int u = 0; u < FAST; ++u) {
                    const int col = (so0 + u) * STRIP + bt;
                    cw[u] = (p.rows && u < nown && col < N) ? ld_relaxed_u64(p.ws_cbox + col) : box_word(0.f);
                }
#pragma unroll
                for (int h = 0; h < 3; ++h)
                    hw[h] = h < nh ? ld_relaxed_u64(p.ws_hbox + ((size_t)blockIdx.x + 1 + h) * STRIP + bt) : box_word(0.f);
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
                                if (bt + b * NBT < hc) v += box_take(dw[a][b], dbox + (size_t)row[a] * MAX_GRID_V2 + bt + b * NBT, err);
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
                const int w = p.xw_world ? N / p.xw_members : 0;
                auto store_y = [&](int col, float yv) {
                    if (p.bias) yv += __ldg(p.bias + col);
                    if (p.xw_world == 0) {
                        if (p.y_is_half) reinterpret_cast<__half *>(p.out)[col] = __float2half_rn(yv);
                        else reinterpret_cast<float *>(p.out)[col] = yv;
                    } else {
                        // local column col of the stacked shard = column j of member m; it lands at [m][rank*w + j] of the
                        // [members][n_full] vector in EVERY rank's arena
                        const int mm = col / w, j = col - mm * w;
                        const unsigned long long e = (unsigned long long)mm * p.xw_nfull + (unsigned long long)p.xw_rank * w + j;
                        for (int pr = 0; pr < p.xw_world; ++pr) {
                            unsigned char *dst = reinterpret_cast<unsigned char *>(__ldg(p.xw_base + pr) + p.xw_out_off);
                            if (p.y_is_half) *reinterpret_cast<__half *>(dst + 2 * e) = __float2half_rn(yv);
                            else *reinterpret_cast<float *>(dst + 4 * e) = yv;
                        }
                    }
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
                        if (p.rows) yv += box_take(cw[u], p.ws_cbox + col, err);
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
                        if (p.rows) yv += box_take(0ull, p.ws_cbox + col, err);
                        if (i == lidx) last_others += yv;
                        else store_y(col, yv);
                    }
                }
                TRACE(15, bt == 0);
#pragma unroll
                for (int h = 0; h < 3; ++h)
                    if (h < nh) last_others += box_take(hw[h], p.ws_hbox + ((size_t)blockIdx.x + 1 + h) * STRIP + bt, err);
                // (a strip that spans many CTAs - narrow column shards: 18 CTAs per strip for a 512-column shard of down_proj - has many such
                //  words: eight reads in flight, not one round trip after the other)
                for (int h0 = 3; h0 < nh; h0 += 8) {
                    unsigned long long mw[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (h0 + k < nh) mw[k] = ld_relaxed_u64(p.ws_hbox + ((size_t)blockIdx.x + 1 + h0 + k) * STRIP + bt);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (h0 + k < nh) last_others += box_take(mw[k], p.ws_hbox + ((size_t)blockIdx.x + 1 + h0 + k) * STRIP + bt, err);
                }
                TRACE(8, bt == 0);
                if (nseg > 0) {
                    const int s = nseg - 1;
                    mbar_wait(bar_u32 + 272 + 8 * (s & 1), (uint32_t)((s >> 1) & 1));
                    if (last_owned) {
                        const uint32_t a = (sacc + 4 * ((s & 1) * STRIP + bt));
                        const float own = lds_f32(a);
                        sts_u32(a, 0u);
                        const int lcol = (s0 + s) * STRIP + bt;
                        if (lcol < N) store_y(lcol, last_others + own);
                    } else {
                        flush(s);
                    }
                }
            }
            
#else
            if (nseg >= 2) {
                mbar_wait(bar_u32 + 272 + 8 * ((nseg - 2) & 1), (uint32_t)(((nseg - 2) >> 1) & 1));
                flush(nseg - 2);
            }
            const long long cb0 = (long long)blockIdx.x * p.chunk, cb1 = min((long long)p.T, cb0 + p.chunk);
            const int so0 = (int)((cb0 + R - 1) / R), so1 = (int)((cb1 + R - 1) / R), nown = so1 - so0;
            const bool last_owned = nseg >= 2 || r0 == 0;  // the last segment's strip starts in this CTA's range (it is strip so1 - 1)
            for (int i = bt; i < nown; i += NBT) {
                const int strip = so0 + i;
                int expect = (int)((((long long)strip + 1) * R - 1) / p.chunk) - (int)blockIdx.x;  // dense CTAs after this one
                if (!SQLLM_CSR_LOCAL && p.rows) {  // every CTA whose share of the CSR rows touches the strip
                    const int c0 = strip * STRIP, c1 = min(N, c0 + STRIP) - 1;
                    expect += c1 / p.csr_rpc - c0 / p.csr_rpc + 1;
                }
                if (p.full_rows) {
                    bool h = false;
                    for (int j = 0; j < p.topX; ++j) {
                        const int c = __ldg(p.fri + j);
                        h |= (c >= 0 && c < N && c / STRIP == strip);
                    }
                    if (h) expect += p.hc;
                }
                if (expect > 0) {
                    int seen;
                    unsigned long long t0 = 0ull, t1;
                    do {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(flags + strip) : "memory");
                        if (seen >= expect) break;
                        __nanosleep(40);
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                        if (t0 == 0ull) t0 = t1;
                        if (t1 - t0 > 2000000000ull) { *reinterpret_cast<volatile int *>(p.ws_cnt + 16) = 1; break; }
                    } while (true);
                    flags[strip] = 0;
                }
            }
            named_bar_sync(3, NBT);  // the polls above are through: the others' contributions to every owned strip are complete
            const float *srowacc = reinterpret_cast<const float *>(sm + OFF_SROW + (SP_ROWS + 1) * 4);
            const bool local_sums = SQLLM_CSR_LOCAL && p.rows && min(N, so1 * STRIP) - min(N, so0 * STRIP) <= SP_ROWS;  // same rule as the sparse warps'
            const int w = p.xw_world ? N / p.xw_members : 0;
            auto store_y = [&](int col, float yv) {
                if (local_sums) yv += srowacc[col - so0 * STRIP];
                if (p.bias) yv += __ldg(p.bias + col);
                if (p.xw_world == 0) {
                    if (p.y_is_half) reinterpret_cast<__half *>(p.out)[col] = __float2half_rn(yv);
                    else reinterpret_cast<float *>(p.out)[col] = yv;
                } else {
                    // local column col of the stacked shard = column j of member m; it lands at [m][rank*w + j] of the
                    // [members][n_full] vector in EVERY rank's arena
                    const int mm = col / w, j = col - mm * w;
                    const unsigned long long e = (unsigned long long)mm * p.xw_nfull + (unsigned long long)p.xw_rank * w + j;
                    for (int pr = 0; pr < p.xw_world; ++pr) {
                        unsigned char *dst = reinterpret_cast<unsigned char *>(__ldg(p.xw_base + pr) + p.xw_out_off);
                        if (p.y_is_half) *reinterpret_cast<__half *>(dst + 2 * e) = __float2half_rn(yv);
                        else *reinterpret_cast<float *>(dst + 4 * e) = yv;
                    }
                }
            };
            // accumulator values first (strips whose own part is already in it: all owned ones but, if owned, the last segment's; and what
            // the others added to that last one - our own part of it never goes through memory), then wait for this CTA's sparse warps
            // (their row sums are needed from here on), then the stores
            const int nearly = last_owned ? nown - 1 : nown;
            float vv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int col = (so0 + u) * STRIP + bt;
                vv[u] = (u < nearly && col < N) ? __ldcg(p.ws_acc + col) : 0.f;
            }
            float pre = 0.f;
            const int lcol = (s0 + nseg - 1) * STRIP + bt;
            if (last_owned && nseg > 0 && lcol < N) pre = __ldcg(p.ws_acc + lcol);
            if (SQLLM_CSR_LOCAL) named_bar_sync(4, (NSPW + NBW) * 32);  // this CTA's own outlier row sums (sparse warps) are complete
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int col = (so0 + u) * STRIP + bt;
                if (u < nearly && col < N) {
                    p.ws_acc[col] = 0.f;
                    store_y(col, vv[u]);
                }
            }
            for (int i = 8; i < nearly; ++i) {  // (more than 8 owned strips: out_features > 75,000 - one at a time)
                const int col = (so0 + i) * STRIP + bt;
                if (col < N) {
                    const float a = __ldcg(p.ws_acc + col);
                    p.ws_acc[col] = 0.f;
                    store_y(col, a);
                }
            }
            if (last_owned && nseg > 0 && lcol < N) p.ws_acc[lcol] = 0.f;
            TRACE(8, bt == 0);
            if (nseg > 0) {
                const int s = nseg - 1;
                mbar_wait(bar_u32 + 272 + 8 * (s & 1), (uint32_t)((s >> 1) & 1));
                if (last_owned) {
                    const float own = lds_f32(sacc + 4 * ((s & 1) * STRIP + bt));
                    if (lcol < N) store_y(lcol, pre + own);
                } else {
                    flush(s);  // a CTA that lies wholly inside a strip started by another one: contribute and announce
                }
            }
#endif
            if (p.xw_world) {
                // exchange: every owning CTA publishes its stores on every rank (system-scope release); CTA 0 (it always owns strip 0)
                // then holds the grid open until every owning CTA of every rank has published on ours: when this grid completes, the
                // local vector is whole.  Counters only grow (expected arrivals so far live next to the flag).
                named_bar_sync(3, NBT);
                if (bt == 0) {
                    const unsigned long long self = __ldg(p.xw_base + p.xw_rank);
                    if (nown > 0) {
                        __threadfence_system();
                        for (int pr = 0; pr < p.xw_world; ++pr)
                            atomicAdd_system(reinterpret_cast<unsigned long long *>(__ldg(p.xw_base + pr) + p.xw_flag_off), 1ull);
                    }
                    if (blockIdx.x == 0) {
                        const unsigned long long target = *reinterpret_cast<volatile unsigned long long *>(self + p.xw_state_off) +
                                                          (unsigned long long)p.xw_world * p.nown_ctas;
                        unsigned long long seen, t0, t1;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                        const bool broken = *reinterpret_cast<volatile unsigned int *>(self + p.xw_err_off) != 0u;
                        if (!broken) do {
                            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(self + p.xw_flag_off) : "memory");
                            if (seen >= target) break;
                            __nanosleep(100);
                            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                            if (t1 - t0 > 2000000000ull) { *reinterpret_cast<volatile unsigned int *>(self + p.xw_err_off) = 1u; break; }
                        } while (true);
                        *reinterpret_cast<volatile unsigned long long *>(self + p.xw_state_off) = target;
                    }
                }
            }
        }
        TRACE(9, bt == 0);
    } else {
        // =========================== consumers ===========================
        const int i16 = lane & 15, jsel = lane >> 4;
        TRACE(2, tid == 0);
        if (tid == 0) pdl_wait();  // everything below reads what the previous kernel may have produced (x)
        named_bar_sync(1, NCT);
        TRACE(3, tid == 0);
        {   // stage x whole, in its own type
            const int n16 = p.K * (XH ? 2 : 4) / 16;
            for (int e = tid; e < n16; e += NCT) cp_async16(xs_u32 + 16 * e, reinterpret_cast<const unsigned char *>(p.x) + 16 * (size_t)e);
            cp_async_commit();
            cp_async_wait_all();
        }
        named_bar_sync(2, NCT + NSPW * 32);  // x visible to consumers and the sparse warp
        TRACE(4, tid == 0);

        Acc A;
#pragma unroll
        for (int t = 0; t < 4; ++t) A.a[t] = 0ull;
#pragma unroll
        for (int t = 0; t < 8; ++t) A.f[t] = 0.f;
        const uint32_t slotb[4] = {(uint32_t)((((0 ^ jsel) << 4) | i16) << 2), (uint32_t)((((1 ^ jsel) << 4) | i16) << 2),
                                   (uint32_t)((((2 ^ jsel) << 4) | i16) << 2), (uint32_t)((((3 ^ jsel) << 4) | i16) << 2)};
        const uint32_t sacc_lane = base + OFF_SACC + (4 * i16) * 4;
        const uint32_t lane_in_stage = (uint32_t)((2 * warp + jsel) * C::UNIT + i16 * 16);
        constexpr int XB = C::XU * (XH ? 2 : 4);  // bytes of x per unit
        constexpr int XBS = SU2 * XB;             // ... per stage

        int slot = 0;
        uint32_t par = 0;
        auto stage_of = [&](int sl) { return (sl < n_lo ? lo_base : hi_base) + sl * C::STAGE + lane_in_stage; };
        auto advance = [&]() { if (++slot == nstage) { slot = 0; par ^= 1u; } };
        for (int seg = 0; seg < nseg; ++seg) {
            const int b = seg & 1;
            // slot addresses of this lane's 4 columns in table b: column (t ^ jsel) -> slot ((t ^ jsel) << 4 | i16); the table base bits above
            // the row field come with them (exact 4-bit: bits 12..15 ride in segc instead, because byte 1 of the address is built by PRMT)
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
            const int mine = 2 * warp < seglen ? (seglen - 2 * warp + SU2 - 1) / SU2 : 0;  // stages in which this warp has a pair

            mbar_wait(bar_u32 + 256 + 8 * b, (uint32_t)((seg >> 1) & 1));  // table b holds this strip
            TRACE(16 + (seg < 7 ? seg : 7), tid == 0);
            // Every consumer warp walks every stage of the segment (the empty barriers count NWC arrivals); a warp whose pair lies past
            // the end of a ragged last stage just releases it.  Two stages per trip, ping-pong: while the gathers of one position run,
            // the words (and x) of the next are already on their way from shared memory into the other register set.
#if SQLLM_V2_PIPE
            Fetch<BITS, XH> F0, F1;
            int k = 0, sl0, sl1;
            mbar_wait(bar_u32 + 8 * slot, par);
            if (0 < mine) fetch2<BITS, XH>(F0, stage_of(slot), xaddr);
            sl0 = slot;
            advance();
            while (true) {
                sl1 = slot;
                if (k + 1 < nstg) {
                    mbar_wait(bar_u32 + 8 * slot, par);
                    if (k + 1 < mine) fetch2<BITS, XH>(F1, stage_of(slot), xaddr + XBS);
                    advance();
                }
                if (k < mine) math2<BITS, MODE, XH>(F0, jsel, l, segc, xaddr, A);
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_u32 + 128 + 8 * sl0);  // stage k: its words went to registers a step ago
                if (k + 1 >= nstg) break;
                sl0 = slot;
                if (k + 2 < nstg) {
                    mbar_wait(bar_u32 + 8 * slot, par);
                    if (k + 2 < mine) fetch2<BITS, XH>(F0, stage_of(slot), xaddr + 2 * XBS);
                    advance();
                }
                if (k + 1 < mine) math2<BITS, MODE, XH>(F1, jsel, l, segc, xaddr + XBS, A);
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_u32 + 128 + 8 * sl1);
                if (k + 2 >= nstg) break;
                k += 2;
                xaddr += 2 * XBS;
            }
#else
            // one stage per trip, no look-ahead: the smallest loop body.  (The two-stage ping-pong variant above hides the word fetch
            // behind the previous position's gathers but doubles the body to ~290 instructions; with 23 warps in five roles on the SM
            // the instruction cache misses ate the gain - ncu: 18 % of the stall samples "no instruction" - see profiles/r02_*.)
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
            {   // this strip's sums: lane (i, j=1) holds column t^1 in slot t - hand it to lane (i, j=0), which adds both into the CTA's
                // shared accumulator of the strip (16 warps x 64 red.shared.add; the builders move it on to global memory)
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
                    const uint32_t a = sacc_lane + b * (STRIP * 4);
                    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(s[0] + v0) : "memory");
                    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a + 4), "f"(s[1] + v1) : "memory");
                    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a + 8), "f"(s[2] + v2) : "memory");
                    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a + 12), "f"(s[3] + v3) : "memory");
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_u32 + 272 + 8 * b);  // done with table b, sums deposited
            }
            TRACE(24 + (seg < 7 ? seg : 7), tid == 0);
        }
        TRACE(6, tid == 0);
    }

    TRACE(11, tid == 0);
}

}  // namespace v2
