// gtb_xtensor.cuh -- the order-exact multi-row Linear on the 5th-generation tensor cores (included by gtb_xrows.cu).
//
// What the reference's Linear demands (gten/ops.h:252-292): per (row, output channel, 32-block, lane l of 4) the INTEGER sum of the
// lane's eight code products, then `acc[l] += float(sum) * (da * dw)` in ascending block order, `(a0+a1)+(a2+a3)` at the end.  The
// fp32 chain cannot go to a tensor core -- but the integer lane sums can, EXACTLY: codes are integers of magnitude <= 127, exact in
// fp16; a lane sum is at most 8 * 127 * 127 < 2^24, so a kind::f16 MMA with fp32 accumulation returns float(sum) bit for bit whatever
// order the hardware adds in.  The four lanes of a block become four N-columns: the staged row of (row r, lane l) keeps lane l's
// eight codes and is zero elsewhere, so D[channel][(r, l)] of ONE K = 32 step (two K16 MMAs) is float(lane sum l).  Each 32-block gets
// its own accumulator (up to eight in flight in TMEM); 16 epilogue warps read every value back (tcgen05.ld) and run the ordered chain:
// per value one FMUL and one FADD, against 8 IDP.4A + unpack + FFMA + FADD in the SIMT kernel (k_xr_gemm).  Measured on the
// micro-benchmark tools/ubench/tmem.cu: 0.12 cycles per (row, channel, block) per SM against 0.28 for the SIMT kernel.
//
// CTA = 128 output channels (UMMA M) x 4*RPT rows (UMMA N = 16*RPT), 21 warps:
//   warps  0..15  epilogue: warp w owns TMEM lanes 32*(w%4).. (channels) and rows (w/4)*RPT.. (columns); RPT*4 fp32 chains per thread
//   warps 16..19  producers: thread = channel: Q4/Q8 codes (cp.async ring, own bytes only) -> fp16 (nibble - 7 / int8) rows of the
//                 128-byte-swizzled A tile; thread = (row, block): staged XBlk codes -> the four masked fp16 rows of the B tile
//   warp  20      issues the MMAs (tcgen05.mma kind::f16, operands from shared memory; the whole warp runs the loop so that the
//                 instruction's uniform operands need no election loop) and owns the TMEM allocation
// One stage = 2 blocks of K (64 fp16 = one 128-byte swizzle row).  The kernel writes the finished fp32 row sums to HBM; the
// re-encode epilogues (xr_epilogue) run in k_xt_epi.
#pragma once

namespace xt {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// a wait that cannot hang the GPU: a broken pipeline traps after ~2 s
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
// the epilogue's wait: between polls the warp sleeps, so that a starved epilogue does not take the issue slots of the producers
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(40);
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// K-major tile, 128-byte swizzle: rows 128 B apart, 8-row groups 1024 B apart (SBO), descriptor version 1, layout 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// four signed bytes -> four fp16 (exact): (b ^ 0x80) = b + 128 as an unsigned byte u; 0x6400 | u is the fp16 1024 + u; minus 1152
__device__ __forceinline__ void s8x4_to_h(uint32_t w, uint32_t& lo, uint32_t& hi) {
    const uint32_t u = w ^ 0x80808080u;
    const __half2 k = __halves2half2(__ushort_as_half((unsigned short)0x6480), __ushort_as_half((unsigned short)0x6480));   // 1152
    uint32_t a = __byte_perm(u, 0x64646464u, 0x4140), b = __byte_perm(u, 0x64646464u, 0x4342);
    const __half2 ha = __hsub2(*reinterpret_cast<__half2*>(&a), k), hb = __hsub2(*reinterpret_cast<__half2*>(&b), k);
    lo = *reinterpret_cast<const uint32_t*>(&ha); hi = *reinterpret_cast<const uint32_t*>(&hb);
}
// four nibble codes held one per byte (0..15) -> four fp16 of (code - 7)
__device__ __forceinline__ void u4x4_to_h(uint32_t u, uint32_t& lo, uint32_t& hi) {
    const __half2 k = __halves2half2(__ushort_as_half((unsigned short)0x6407), __ushort_as_half((unsigned short)0x6407));   // 1031
    uint32_t a = __byte_perm(u, 0x64646464u, 0x4140), b = __byte_perm(u, 0x64646464u, 0x4342);
    const __half2 ha = __hsub2(*reinterpret_cast<__half2*>(&a), k), hb = __hsub2(*reinterpret_cast<__half2*>(&b), k);
    lo = *reinterpret_cast<const uint32_t*>(&ha); hi = *reinterpret_cast<const uint32_t*>(&hb);
}

}  // namespace xt

constexpr int XT_BM = 128;                 // output channels per CTA = UMMA M
constexpr int XT_RING = 8;                 // raw weight bytes in flight per channel, in stages
constexpr int XT_EPI_WARPS = 16, XT_PROD_WARPS = 4;
constexpr int XT_NT = (XT_EPI_WARPS + XT_PROD_WARPS + 1) * 32;  // 21 warps: 80 registers per thread

struct XtGemmArgs {
    const XBlk* act; int nb;           // staged input rows [row][nb]
    const uint4* wd; const uint16_t* ws; int N;
    int row0, n_rows;                  // rows [row0, row0 + n_rows); blockIdx.y = group of 4*RPT rows
    float* out; int ldo; int out_row_sub;   // out[(row - out_row_sub) * ldo + channel] = the Linear's fp32 result
    long long* dbg;                    // option "xr_trace": cycle counters of CTA (0, 0), accumulated over launches
};

template <int WT, int RPT>
struct XtCfg {
    static constexpr int N = 16 * RPT;                                   // UMMA N = rows x 4 lanes
    static constexpr int ROWS = 4 * RPT;
    static constexpr int RAWB = (WT == DT_Q4) ? 32 : 64;                 // raw weight bytes per channel per stage
    static constexpr uint32_t A_BYTES = XT_BM * 128, B_BYTES = N * 128, AD_BYTES = 1024;
    static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES + AD_BYTES;
    static constexpr uint32_t RAW_BYTES = XT_RING * XT_BM * RAWB + XT_RING * 2 * ROWS * 48;   // weight bytes + staged activation blocks in flight
    static constexpr int STAGES = (RPT >= 8) ? 4 : ((WT == DT_Q8 && RPT == 4) ? 5 : 6);   // stages of 2 blocks
    static constexpr int NACC = (512 / N < 8) ? 512 / N : 8;             // accumulators in TMEM: one per block in flight
    static constexpr uint32_t SMEM = STAGES * STAGE_BYTES + RAW_BYTES + 512 + 1024;
    static constexpr uint32_t TMEM_COLS = NACC * N;
    static_assert(SMEM <= 227 * 1024, "stage ring + raw rings exceed the shared memory of an SM");
    // kind::f16: D = fp32 (bits 4-5 = 1), A = B = fp16 (0), K-major both, N >> 3 at bits 17-22, M >> 4 at bits 24-28
    static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(XT_BM >> 4) << 24);
};

template <int WT, int RPT>
__global__ void __launch_bounds__(XT_NT, 1) k_xt_gemm(XtGemmArgs a) {
    using Cfg = XtCfg<WT, RPT>;
    constexpr int S = Cfg::STAGES, NACC = Cfg::NACC;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (xt::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* raw_ring = smem + S * Cfg::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(raw_ring + Cfg::RAW_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 2 * NACC);
    const uint32_t smem_base = xt::smem_u32(smem), bar_base = xt::smem_u32(bars);
    auto bar_full = [&](int s) { return bar_base + 8u * s; };
    auto bar_empty = [&](int s) { return bar_base + 8u * (S + s); };
    auto bar_tfull = [&](int x) { return bar_base + 8u * (2 * S + x); };
    auto bar_tempty = [&](int x) { return bar_base + 8u * (2 * S + NACC + x); };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ch0 = blockIdx.x * XT_BM;
    const int rbase = a.row0 + blockIdx.y * Cfg::ROWS, rend = a.row0 + a.n_rows;
    const int nb = a.nb, nst = nb / 2;

    // the B tiles' masked-out chunks and the scales of rows past the end stay zero for the whole kernel
    for (uint32_t i = tid; i < S * Cfg::STAGE_BYTES / 16; i += XT_NT) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        for (int s = 0; s < S; s++) { xt::mbar_init(bar_full(s), XT_PROD_WARPS * 32); xt::mbar_init(bar_empty(s), 1 + XT_EPI_WARPS); }
        for (int x = 0; x < NACC; x++) { xt::mbar_init(bar_tfull(x), 1); xt::mbar_init(bar_tempty(x), XT_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == XT_EPI_WARPS + XT_PROD_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(xt::smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    xt::proxy_fence();
    xt::tc_fence_before();
    __syncthreads();
    xt::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < XT_EPI_WARPS) {
        // ------------------------------------------------------------ epilogue: the ordered fp32 chains
        const int lg = warp & 3, cg = warp >> 2;
        const int ch = ch0 + lg * 32 + lane;
        const uint16_t* wsp = a.ws + (size_t)min(ch, a.N - 1) * nb;
        float acc[RPT][4];
#pragma unroll
        for (int r = 0; r < RPT; r++)
#pragma unroll
            for (int l = 0; l < 4; l++) acc[r][l] = 0.0f;
        uint32_t pair = __ldg(reinterpret_cast<const uint32_t*>(wsp));      // the channel's scales of the stage's two blocks
        const uint32_t tbase = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(cg * RPT * 4);
        const bool tr = a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0;
        long long c_wait = 0, c_ld = 0;
        const long long c_t0 = tr ? clock64() : 0;
        for (int j = 0; j < nst; j++) {
            const int s = j % S;
            const unsigned char* st = smem + s * Cfg::STAGE_BYTES;
            const float* adp = reinterpret_cast<const float*>(st + Cfg::A_BYTES + Cfg::B_BYTES);
            const uint32_t pair_next = (j + 1 < nst) ? __ldg(reinterpret_cast<const uint32_t*>(wsp + 2 * (j + 1))) : 0u;
#pragma unroll
            for (int blk = 0; blk < 2; blk++) {
                const float dw = __half2float(__ushort_as_half((unsigned short)(blk ? (pair >> 16) : (pair & 0xffffu))));
                const int g = 2 * j + blk, ai = g & (NACC - 1);
                long long c0 = tr ? clock64() : 0;
                xt::mbar_wait_backoff(bar_tfull(ai), (g / NACC) & 1);
                xt::tc_fence_after();
                if (tr) { const long long c1 = clock64(); c_wait += c1 - c0; c0 = c1; }
                const uint32_t taddr = tbase + (uint32_t)(ai * Cfg::N);
                constexpr int NLD = (RPT >= 4) ? RPT / 4 : 1;            // loads of 16 columns (4 rows); RPT = 2: one load of 8
#pragma unroll
                for (int h = 0; h < NLD; h++) {
                    uint32_t d[16];
                    if (RPT >= 4) {
                        xt::tmem_ld16(taddr + h * 16, d);
                    } else {
                        uint32_t d8[8];
                        xt::tmem_ld8(taddr, d8);
#pragma unroll
                        for (int i = 0; i < 8; i++) d[i] = d8[i];
                    }
                    if (tr) { const long long c1 = clock64(); c_ld += c1 - c0; }
                    if (h == NLD - 1) {                                    // this accumulator may be overwritten by block g + NACC
                        xt::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) xt::mbar_arrive(bar_tempty(ai));
                    }
                    constexpr int RL = (RPT >= 4) ? 4 : RPT;
#pragma unroll
                    for (int r = 0; r < RL; r++) {
                        const int rr = h * 4 + r;
                        // (float)lane_sum * (da * dw), added in block order (gten/ops.h:282-287)
                        const float sdw = __fmul_rn(adp[blk * Cfg::ROWS + cg * RPT + rr], dw);
#pragma unroll
                        for (int l = 0; l < 4; l++) acc[rr][l] = __fadd_rn(acc[rr][l], __fmul_rn(__uint_as_float(d[4 * r + l]), sdw));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) xt::mbar_arrive(bar_empty(s));                  // done with the stage's activation scales
            pair = pair_next;
        }
        if (tr) { a.dbg[8] += clock64() - c_t0; a.dbg[9] += c_wait; a.dbg[10] += c_ld; a.dbg[12] += nst; a.dbg[0] += 1; }
#pragma unroll
        for (int r = 0; r < RPT; r++) {
            const int row = rbase + cg * RPT + r;
            if (row < rend && ch < a.N)
                a.out[(size_t)(row - a.out_row_sub) * a.ldo + ch] = __fadd_rn(__fadd_rn(acc[r][0], acc[r][1]), __fadd_rn(acc[r][2], acc[r][3]));
        }
    } else if (warp < XT_EPI_WARPS + XT_PROD_WARPS) {
        // ------------------------------------------------------------ producers
        const int pt = tid - XT_EPI_WARPS * 32;                            // 0..127: channel of the A tile
        constexpr int RAWB = Cfg::RAWB;
        const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(a.wd) + (size_t)min(ch0 + pt, a.N - 1) * nb * (RAWB / 2);
        unsigned char* myraw = raw_ring + (size_t)pt * RAWB;
        // B tile: thread = (local row, block of the stage); 2 * ROWS pairs.  The staged activation block (codes + scale, 48 of the
        // record's 64 bytes) rides the same cp.async ring as the weights: XT_RING stages ahead of its use
        constexpr int NPAIR = 2 * Cfg::ROWS;
        const bool has_pair = pt < NPAIR;
        const int prow = pt >> 1, pblk = pt & 1;
        const bool row_ok = has_pair && (rbase + prow) < rend;
        const XBlk* asrc = a.act + (size_t)(row_ok ? rbase + prow : a.row0) * nb + pblk;
        unsigned char* bring = raw_ring + (size_t)XT_RING * XT_BM * RAWB + (size_t)(has_pair ? pt : 0) * 48;
        auto issue_raw = [&](int j) {
            if (j < nst) {
                unsigned char* dst = myraw + (size_t)(j % XT_RING) * XT_BM * RAWB;
#pragma unroll
                for (int i = 0; i < RAWB / 16; i++) cp_async16(dst + 16 * i, wsrc + (size_t)j * RAWB + 16 * i, true);
                if (row_ok) {
                    const unsigned char* src = reinterpret_cast<const unsigned char*>(asrc + 2 * j);
                    unsigned char* bd = bring + (size_t)(j % XT_RING) * NPAIR * 48;
                    cp_async16(bd, src, true); cp_async16(bd + 16, src + 16, true); cp_async16(bd + 32, src + 48, true);
                }
            }
            cp_async_commit();
        };
#pragma unroll
        for (int j = 0; j < XT_RING - 1; j++) issue_raw(j);
        const bool tr = a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && pt == 32;
        long long c_empty = 0, c_cp = 0, c_work = 0;
        const long long c_t0 = tr ? clock64() : 0;
        for (int j = 0; j < nst; j++) {
            const int s = j % S;
            issue_raw(j + XT_RING - 1);
            long long c0 = tr ? clock64() : 0;
            xt::mbar_wait(bar_empty(s), ((j / S) & 1) ^ 1);
            if (tr) { const long long c1 = clock64(); c_empty += c1 - c0; c0 = c1; }
            unsigned char* st = smem + s * Cfg::STAGE_BYTES;
            // ---- A: this channel's two blocks -> fp16, chunk (blk * 4 + l) of the 128-byte row holds lane l: x0..x3, y0..y3
            cp_async_wait<XT_RING - 1>();
            if (tr) { const long long c1 = clock64(); c_cp += c1 - c0; c0 = c1; }
            {
                const uint4* rw = reinterpret_cast<const uint4*>(myraw + (size_t)(j % XT_RING) * XT_BM * RAWB);
                unsigned char* arow = st + (pt >> 3) * 1024 + (pt & 7) * 128;
#pragma unroll
                for (int blk = 0; blk < 2; blk++) {
                    uint32_t wx[4], wy[4];
                    if (WT == DT_Q4) {
                        const uint4 q = rw[blk];
                        const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int l = 0; l < 4; l++) { wx[l] = (qq[l] >> 4) & 0x0f0f0f0fu; wy[l] = qq[l] & 0x0f0f0f0fu; }
                    } else {
                        const uint4 x = rw[2 * blk], y = rw[2 * blk + 1];
                        wx[0] = x.x; wx[1] = x.y; wx[2] = x.z; wx[3] = x.w; wy[0] = y.x; wy[1] = y.y; wy[2] = y.z; wy[3] = y.w;
                    }
#pragma unroll
                    for (int l = 0; l < 4; l++) {
                        uint4 o;
                        if (WT == DT_Q4) { xt::u4x4_to_h(wx[l], o.x, o.y); xt::u4x4_to_h(wy[l], o.z, o.w); }
                        else { xt::s8x4_to_h(wx[l], o.x, o.y); xt::s8x4_to_h(wy[l], o.z, o.w); }
                        *reinterpret_cast<uint4*>(arow + (((blk * 4 + l) ^ (pt & 7)) << 4)) = o;
                    }
                }
            }
            // ---- B: the four masked rows (prow, l) of block pblk
            if (row_ok) {
                const unsigned char* bs = bring + (size_t)(j % XT_RING) * NPAIR * 48;
                const uint4 cx = *reinterpret_cast<const uint4*>(bs), cy = *reinterpret_cast<const uint4*>(bs + 16);
                const float cd = *reinterpret_cast<const float*>(bs + 32);
                const uint32_t ax[4] = {cx.x, cx.y, cx.z, cx.w}, ay[4] = {cy.x, cy.y, cy.z, cy.w};
                unsigned char* bt = st + Cfg::A_BYTES;
#pragma unroll
                for (int l = 0; l < 4; l++) {
                    const int nr = prow * 4 + l;
                    uint4 o;
                    xt::s8x4_to_h(ax[l], o.x, o.y); xt::s8x4_to_h(ay[l], o.z, o.w);
                    *reinterpret_cast<uint4*>(bt + (nr >> 3) * 1024 + (nr & 7) * 128 + (((pblk * 4 + l) ^ (nr & 7)) << 4)) = o;
                }
                reinterpret_cast<float*>(st + Cfg::A_BYTES + Cfg::B_BYTES)[pblk * Cfg::ROWS + prow] = cd;
            }
            xt::proxy_fence();                                             // generic stores -> visible to the tensor core's reads
            xt::mbar_arrive(bar_full(s));
            if (tr) { const long long c1 = clock64(); c_work += c1 - c0; c0 = c1; }
        }
        if (tr) {
            long long* d = a.dbg + 24;
            d[0] += clock64() - c_t0; d[1] += c_empty; d[2] += c_cp; d[3] += c_work;
        }
        cp_async_wait<0>();
    } else {
        // ------------------------------------------------------------ MMA issuer: one accumulator per block, each a fresh sum (two K16 steps)
        const bool tr = a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
        long long c_full = 0, c_tempty = 0;
        const long long c_t0 = tr ? clock64() : 0;
        for (int j = 0; j < nst; j++) {
            const int s = j % S;
            long long c0 = tr ? clock64() : 0;
            xt::mbar_wait(bar_full(s), (j / S) & 1);
            xt::tc_fence_after();
            if (tr) c_full += clock64() - c0;
            const uint32_t sa = smem_base + s * Cfg::STAGE_BYTES, sb = sa + Cfg::A_BYTES;
#pragma unroll
            for (int blk = 0; blk < 2; blk++) {
                const int g = 2 * j + blk, ai = g & (NACC - 1);
                if (tr) c0 = clock64();
                xt::mbar_wait(bar_tempty(ai), ((g / NACC) & 1) ^ 1);       // the epilogue has read block g - NACC out of this accumulator
                xt::tc_fence_after();
                if (tr) c_tempty += clock64() - c0;
                if (lane == 0) {
                    const uint32_t d_tmem = tmem + (uint32_t)(ai * Cfg::N);
                    xt::umma_f16(d_tmem, xt::umma_desc_sw128(sa + blk * 64), xt::umma_desc_sw128(sb + blk * 64), Cfg::IDESC, 0u);
                    xt::umma_f16(d_tmem, xt::umma_desc_sw128(sa + blk * 64 + 32), xt::umma_desc_sw128(sb + blk * 64 + 32), Cfg::IDESC, 1u);
                    xt::umma_commit(bar_tfull(ai));
                }
                __syncwarp();
            }
            if (lane == 0) xt::umma_commit(bar_empty(s));                  // the stage's tiles are free once these MMAs have read them
            __syncwarp();
        }
        if (tr) { a.dbg[16] += clock64() - c_t0; a.dbg[20] += c_full; a.dbg[21] += c_tempty; }
    }
    xt::tc_fence_before();
    __syncthreads();
    if (warp == XT_EPI_WARPS + XT_PROD_WARPS)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
}

// The re-encode epilogues over the finished rows: one warp per (row, tile n) -- tile = 64 output columns (gate|up: 32 channels)
template <int EPI>
__global__ void __launch_bounds__(256) k_xt_epi(XrGemmArgs a, const float* __restrict__ raw, int ld) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n = blockIdx.x, row = a.row0 + blockIdx.y * 8 + wid;
    if (row >= a.row0 + a.n_rows) return;                                  // warp-uniform
    const XrRow rw = a.rows[row];
    if (rw.slot < 0) return;
    const float* src = raw + (size_t)(row - ((EPI == XEPI_HEAD) ? a.row0 : 0)) * ld;
    float v[2];
    if (EPI == XEPI_SILU) {
        v[0] = src[32 * n + lane]; v[1] = src[a.up_off + 32 * n + lane];
    } else {
        const int c0 = 64 * n + lane, c1 = c0 + 32;
        v[0] = (c0 < a.N) ? src[c0] : 0.0f; v[1] = (c1 < a.N) ? src[c1] : 0.0f;
    }
    xr_epilogue<EPI>(a, row, rw, n, lane, v, false);
}
